// K2s: sequence-first one-byte tokens -- the reference's DEFAULT layout (batch_first=False, src/tokenize.cpp:82-98;
// stores ptr[s * B + i], src/tokenize.h:435-439) -- persistent and warp-specialised (sm_100a).
//
// Tile = 128 sequences x 128 positions = 16 KiB of the (padlen, nseq) output: 128 rows of 128 contiguous bytes.
//
//   2 producer warps per tile: resolve the 128 sequences (offsets, length, source alignment), then stage the 16-byte
//                   aligned window of residues that lies under the tile's 128 columns of every sequence into the next free
//                   stage of a ring with per-lane 16-byte asynchronous copies (cp.async, eight lanes per sequence), whose
//                   completion is counted on the stage's `full` mbarrier (cp.async.mbarrier.arrive).  One bulk (TMA) copy
//                   per sequence was measured first: 128 serialised copy issues per 16 KiB tile made the producer the
//                   bottleneck (178 vs 100 us on C2x4, profiles/r02n).  Tiles are handed out position-fastest, so
//                   neighbouring windows of a sequence are fetched back to back and share their L2 sectors.
//   8 consumer warps per tile, two phases around a token tile in shared memory:
//     phase 1  lane = 16 consecutive columns of one sequence: realign (five LDS.32 + four funnel shifts), 16 LUT look-ups,
//              BOS / EOS / PAD from the mask tables on boundary vectors, one STS.128 into tile[seq][pos], 16-byte chunks
//              XOR-swizzled by the sequence group;
//     phase 2  thread = 16 sequences x 4 positions: 16 conflict-free LDS.32, four 4x4 byte transposes in registers (PRMT),
//              one st.global.cs.v4 per position; eight neighbouring lanes write 128 contiguous bytes of an output row.
//
// What this replaces (K2t, seqfirst_tok8_kernel in bsq_kernels.cu): one CTA per tile, whose LUT / mask-table set-up,
// offsets round trip, address arithmetic of the staging copies and their exposed latency were paid per 16 KiB of output
// (ncu r01e: 133 warp-instructions per 16 bytes per lane, issue-active 59-63 %, 70 % of the warps resident; 0.64-0.75 of
// the copy peak).  Here all of that is either done once per CTA or moved to the producer warp, which runs a ring of
// stages ahead of the consumers.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "bsq_internal.h"
#include "bsq_kernels.cuh"

namespace bsq {

namespace {

constexpr unsigned kWaitBackoffNs = 64;

constexpr int kSfSeqs = 128, kSfPos = 128, kSfPitch = kSfPos + 32;  // stage pitch: a 128-column window at any alignment + the second LDS.128 of its last chunk
constexpr int kSfConsumerWarps = 8;
constexpr int kSfConsumers = kSfConsumerWarps * 32;
constexpr int kSfProducerWarps = 2;
constexpr int kSfProducers = kSfProducerWarps * 32;
constexpr int kSfThreads = kSfConsumers + kSfProducers;
constexpr int kSfStageBytes = kSfSeqs * kSfPitch;  // 18432
constexpr int kSfTileBytes = kSfSeqs * kSfPos;     // 16384

__device__ __forceinline__ uint32_t sf_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void sf_bar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sf_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sf_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sf_wait(uint32_t bar, uint32_t parity) {
    // A warp that polls in a tight loop takes issue slots from the warps that have work (ncu r02n: 40 % of the
    // executed instructions were polls); after a failed first try it backs off with nanosleep between tries.
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra BSQ_SF_WAIT_DONE_%=;\n"
        "BSQ_SF_WAIT_%=:\n"
        "nanosleep.u32 %2;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra BSQ_SF_WAIT_%=;\n"
        "BSQ_SF_WAIT_DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity), "r"(kWaitBackoffNs)
        : "memory");
}
__device__ __forceinline__ void sf_cp16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// arrival on `bar` once every cp.async this thread has issued so far has landed (counts against the barrier's expected
// arrivals: .noinc)
__device__ __forceinline__ void sf_cp_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// 4x4 byte transpose: x[r] holds 4 consecutive positions of sequence r; y[j] gets position j of the 4 sequences.
__device__ __forceinline__ void sf_transpose4x4(const uint32_t x[4], uint32_t y[4]) {
    const uint32_t t0 = __byte_perm(x[0], x[1], 0x5140), t1 = __byte_perm(x[2], x[3], 0x5140);
    const uint32_t t2 = __byte_perm(x[0], x[1], 0x7362), t3 = __byte_perm(x[2], x[3], 0x7362);
    y[0] = __byte_perm(t0, t1, 0x5410);
    y[1] = __byte_perm(t0, t1, 0x7632);
    y[2] = __byte_perm(t2, t3, 0x5410);
    y[3] = __byte_perm(t2, t3, 0x7632);
}

struct SfParams {
    const uint8_t *bytes;
    const int64_t *offs;
    uint8_t *out;
    int64_t nseq;     // sequences of this launch (columns written)
    int64_t ld;       // batch extent of the whole output array (row pitch in bytes)
    int64_t ntiles;
    int padlen;
    int ptiles;       // position tiles per sequence group
    uint32_t div_mul, div_shift;  // tile / ptiles for 0 <= tile < 2^31
    unsigned int *ctr;            // dynamic tile counter (see bsq_span.cu), or nullptr: tile += gridDim.x
};

// The stage is one buffer of 128 windows used as two halves of 64 sequences, each with its own full / empty barrier: while
// the consumers translate one half, the producers refill the other, and during phase 2 (which only touches the token tile)
// both halves of the next tile arrive.  That is double buffering at the shared-memory footprint of a single stage
// (34 KiB: five CTAs per SM; with two whole stages only three fitted, and the consumers are latency-bound: 115 vs 99 us).
template <int MINB>
__global__ void __launch_bounds__(kSfThreads, MINB)
tokenize_seqfirst_kernel(const SfParams q, const LutParam lutp, const Specials sp) {
    extern __shared__ __align__(128) uint8_t dyn[];  // stage[128][144], then the token tile [128][128]
    __shared__ __align__(16) uint8_t lut[256];
    __shared__ TailTab tab;
    __shared__ __align__(8) uint64_t full[2], empty[2];
    __shared__ __align__(16) int4 rinfo[kSfSeqs];  // per sequence: n = bos + len, A & 15, A & ~15 (A = global address of column p0)
    __shared__ int4 hdr[2];                        // per half: first sequence (lo, hi), first column, sequences in the tile (<0: end)
    __shared__ unsigned int pub_tile[2];           // dynamic tile order: the counter's answer, for both producer warps

    asm volatile("griddepcontrol.launch_dependents;");
    load_lut(lut, lutp);
    init_tailtab(tab, sp);
    if (threadIdx.x == 128) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            sf_bar_init(full + h, 2 * kSfProducers);  // per producer thread: its copies have landed + its table entries are written
            sf_bar_init(empty + h, kSfConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t *const tile = dyn + kSfStageBytes;

    if (warp >= kSfConsumerWarps) {
        // ------------------------------- producers -------------------------------
        const int pwarp = warp - kSfConsumerWarps;  // 32 of the 64 sequences of every half
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const int maxlen = max(q.padlen - sp.bos - sp.eos, 0);
        const uint32_t stage = sf_u32(dyn);
        const int qv = lane & 7;
        int64_t t = blockIdx.x;
        for (uint32_t it = 0;; ++it) {
            const uint32_t ph = it & 1u;
            // the tile after this one: asked for now (producer warp 0, lane 0), needed at the end of the iteration
            unsigned int asked = 0;
            if (q.ctr != nullptr && pwarp == 0 && lane == 0) asked = atomicAdd(q.ctr, 1u);
            // tiles are numbered position-fastest within a group of 128 sequences
            const uint32_t g = q.ptiles == 1 ? static_cast<uint32_t>(t) : (__umulhi(static_cast<uint32_t>(t), q.div_mul) >> q.div_shift);
            const int p0 = static_cast<int>(static_cast<uint32_t>(t) - g * static_cast<uint32_t>(q.ptiles)) * kSfPos;
            const int64_t i0 = static_cast<int64_t>(g) * kSfSeqs;
            const int ns = static_cast<int>(min(static_cast<int64_t>(kSfSeqs), q.nseq - i0));
            // step A: lane resolves one sequence of each half (offsets in flight while the half is waited for)
            int4 ri[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int il = 64 * h + 32 * pwarp + lane;
                ri[h] = make_int4(0, 0, 0, 0);
                if (il < ns) {
                    const int64_t start = __ldg(q.offs + i0 + il);
                    const int len = static_cast<int>(min(max(__ldg(q.offs + i0 + il + 1) - start, int64_t(0)), static_cast<int64_t>(maxlen)));
                    const uintptr_t a = reinterpret_cast<uintptr_t>(q.bytes + (start - sp.bos + p0));  // source address of column p0
                    ri[h] = make_int4(sp.bos + len, static_cast<int>(a & 15u), static_cast<int>(static_cast<uint32_t>(a & ~uintptr_t(15))),
                                      static_cast<int>(static_cast<uint32_t>(a >> 32)));
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (it > 0) sf_wait(sf_u32(empty + h), ph ^ 1u);  // the consumers are done with this half
                rinfo[64 * h + 32 * pwarp + lane] = ri[h];
                if (pwarp == 0 && lane == 0) hdr[h] = make_int4(static_cast<int>(i0 & 0xffffffffll), static_cast<int>(i0 >> 32), p0, ns);
                __syncwarp();
                // step B: eight lanes per sequence copy the 16-byte chunks of its window that hold residues of this tile
                const uint32_t bar = sf_u32(full + h);
#pragma unroll 4
                for (int u = 0; u < 8; ++u) {
                    const int il = 64 * h + 32 * pwarp + 4 * u + (lane >> 3);
                    const int4 r = rinfo[il];
                    // stage offsets [olo, ohi) of the row hold the residues under this tile (columns max(bos, p0) .. min(n, p0 + 128))
                    const int olo = r.y + max(sp.bos - p0, 0), ohi = r.y + min(r.x - p0, kSfPos);
                    const uint8_t *src = reinterpret_cast<const uint8_t *>((static_cast<uint64_t>(static_cast<uint32_t>(r.w)) << 32) | static_cast<uint32_t>(r.z));
                    const uint32_t dst = stage + static_cast<uint32_t>(il * kSfPitch) + 16u * qv;
                    if (16 * qv < ohi && 16 * qv + 16 > olo) sf_cp16(dst, src + 16 * qv);
                    if (qv == 0 && ohi > kSfPos) sf_cp16(dst + kSfPos, src + kSfPos);
                }
                sf_cp_arrive(bar);  // when this thread's copies have landed
                sf_arrive(bar);     // its table entries are written (release)
            }
            // next tile
            int64_t tn;
            if (q.ctr == nullptr) {
                tn = t + gridDim.x;
            } else {
                if (pwarp == 0 && lane == 0) pub_tile[it & 1] = asked;
                // both producer warps agree through a named barrier (64 threads); the slot alternates, so the next
                // iteration's write cannot overtake this iteration's reads
                asm volatile("bar.sync 2, %0;" ::"n"(kSfProducers) : "memory");
                tn = static_cast<int64_t>(gridDim.x) + pub_tile[it & 1];
            }
            if (tn >= q.ntiles) {
                // end of work for this CTA: a first half whose header says so
                sf_wait(sf_u32(empty + 0), ph);
                if (pwarp == 0 && lane == 0) hdr[0] = make_int4(0, 0, 0, -1);
                __syncwarp();
                sf_arrive(sf_u32(full + 0));
                sf_arrive(sf_u32(full + 0));
                break;
            }
            t = tn;
        }
        if (q.ctr != nullptr && pwarp == 0 && lane == 0) {
            __threadfence();
            if (atomicAdd(q.ctr + 1, 1u) == gridDim.x - 1) {
                q.ctr[0] = 0u;
                q.ctr[1] = 0u;
                __threadfence();
            }
        }
        return;
    }

    // ------------------------------- consumers -------------------------------
    const int tid = threadIdx.x;
    const int qv = tid & 7;  // 16-byte column chunk of phase 1
    const uint4 padv = make_uint4(sp.pad_w, sp.pad_w, sp.pad_w, sp.pad_w);
    for (uint32_t it = 0;; ++it) {
        const uint32_t ph = it & 1u;
        int64_t i0 = 0;
        int p0 = 0, ns = 0;
        // ---- phase 1: tile[seq][pos] <- codes, one half of the stage after the other ----
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            sf_wait(sf_u32(full + h), ph);
            const int4 hd = hdr[h];
            if (h == 0) {
                if (hd.w < 0) return;  // no more tiles for this CTA
                i0 = (static_cast<int64_t>(hd.y) << 32) | static_cast<uint32_t>(hd.x);
                p0 = hd.z;
                ns = hd.w;
            }
            const int c0 = p0 + 16 * qv;
            const bool bos_here = sp.bos && c0 == 0;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int il = 64 * h + (tid >> 3) + 32 * u;
                const int4 ri = rinfo[il];
                const int n = ri.x;
                uint4 codes = padv;
                if (c0 < n + sp.eos) {
                    uint32_t t[4] = {0u, 0u, 0u, 0u};
                    if (c0 < n) {
                        // two aligned LDS.128 (conflict-free: the 8 lanes of a sequence read 128 contiguous bytes) and a
                        // per-lane word select, instead of five LDS.32 at the lane's own word offset (4-way bank conflicts:
                        // 20 shared-memory wavefronts per vector against 8, in a kernel bounded by them)
                        const uint32_t o = static_cast<uint32_t>(ri.y) + 16u * qv;
                        const uint4 v0 = *reinterpret_cast<const uint4 *>(dyn + il * kSfPitch + (o & ~15u));
                        const uint4 v1 = *reinterpret_cast<const uint4 *>(dyn + il * kSfPitch + (o & ~15u) + 16);
                        const bool q1 = (o & 4u) != 0, q2 = (o & 8u) != 0;
                        const uint32_t a0 = q1 ? v0.y : v0.x, a1 = q1 ? v0.z : v0.y, a2 = q1 ? v0.w : v0.z, a3 = q1 ? v1.x : v0.w;
                        const uint32_t a4 = q1 ? v1.y : v1.x, a5 = q1 ? v1.z : v1.y, a6 = q1 ? v1.w : v1.z;
                        const uint32_t w0 = q2 ? a2 : a0, w1 = q2 ? a3 : a1, w2 = q2 ? a4 : a2, w3 = q2 ? a5 : a3, w4 = q2 ? a6 : a4;
                        const uint32_t sh = (o & 3u) * 8u;
                        t[0] = translate4(__funnelshift_r(w0, w1, sh), lut);
                        t[1] = translate4(__funnelshift_r(w1, w2, sh), lut);
                        t[2] = translate4(__funnelshift_r(w2, w3, sh), lut);
                        t[3] = translate4(__funnelshift_r(w3, w4, sh), lut);
                    }
                    if (bos_here) t[0] = __byte_perm(t[0], sp.bos_w, 0x3214);
                    if (c0 + 16 > n) {  // the row ends inside this chunk
                        const uint4 m = tab.m[n - c0], f = tab.f[n - c0];
                        t[0] = (t[0] & m.x) | f.x; t[1] = (t[1] & m.y) | f.y;
                        t[2] = (t[2] & m.z) | f.z; t[3] = (t[3] & m.w) | f.w;
                    }
                    codes = make_uint4(t[0], t[1], t[2], t[3]);
                }
                // 16-byte chunks XOR-swizzled by the sequence group, so that phase 2's column-of-words reads
                // (16 sequences apart) hit 32 distinct banks
                *reinterpret_cast<uint4 *>(tile + il * kSfPos + 16 * (qv ^ ((il >> 4) & 7))) = codes;
            }
            // this half of the stage (and its table entries) has been read: hand it back
            __syncwarp();
            if (lane == 0) sf_arrive(sf_u32(empty + h));
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kSfConsumers) : "memory");  // the whole token tile is written

        // ---- phase 2: thread = 16 sequences (group A) x 4 positions (word pw) ----
        const int A = lane & 7, pw = 4 * warp + (lane >> 3);
        const uint32_t *t32 = reinterpret_cast<const uint32_t *>(tile) + 16 * A * (kSfPos / 4) + (pw ^ (4 * A));
        uint32_t y[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            uint32_t x[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = t32[(4 * a + k) * (kSfPos / 4)];
            sf_transpose4x4(x, y[a]);
        }
        const int pos = p0 + 4 * pw;
        uint8_t *o = q.out + (static_cast<int64_t>(pos) * q.ld + i0 + 16 * A);
        if (ns == kSfSeqs && p0 + kSfPos <= q.padlen) {
#pragma unroll
            for (int j = 0; j < 4; ++j, o += q.ld) __stcs(reinterpret_cast<uint4 *>(o), make_uint4(y[0][j], y[1][j], y[2][j], y[3][j]));
        } else {
            const int nvalid = ns - 16 * A;  // sequences of this group that exist
#pragma unroll
            for (int j = 0; j < 4; ++j, o += q.ld) {
                if (pos + j >= q.padlen || nvalid <= 0) continue;
                if (nvalid >= 16) {
                    __stcs(reinterpret_cast<uint4 *>(o), make_uint4(y[0][j], y[1][j], y[2][j], y[3][j]));
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (k < nvalid) o[k] = static_cast<uint8_t>(y[k >> 2][j] >> (8 * (k & 3)));
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kSfConsumers) : "memory");  // the token tile is rewritten by the next tile's phase 1
    }
}

int sf_env(const char *name, int dflt) {
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

}  // namespace

bool seqfirst_kernel_applicable() {
    static const bool tune = sf_env("BSQ_TUNE", 0) != 0;
    static bool on = sf_env("BSQ_SF2", 1) != 0;
    if (tune) on = sf_env("BSQ_SF2", 1) != 0;
    return on;
}

// K2s over `nseq` sequences; out is the (padlen, ld) array's first column of this range.  The caller guarantees
// one-byte codes (every alphabet but BYTES), a 16-byte aligned `out` and ld % 16 == 0.
int launch_tokenize_seqfirst(int device, cudaStream_t st, const SeqView &v, int64_t nseq, int64_t ld, int64_t padlen, const Prepared &p,
                             uint8_t *out, bool pdl_allowed) {
    static const bool tune = sf_env("BSQ_TUNE", 0) != 0;
    static int minb = 0, ctas = 0, dynamic = 1;
    if (minb == 0 || tune) {
        minb = std::min(6, std::max(4, sf_env("BSQ_SF2_MINB", 5)));  // launch bound: 4 -> 48, 5 -> 40, 6 -> 32 registers
        ctas = std::max(1, sf_env("BSQ_SF2_CTAS", minb));
        dynamic = sf_env("BSQ_SF2_DYN", 1);
    }
    const int64_t groups = (nseq + kSfSeqs - 1) / kSfSeqs, ptiles = (padlen + kSfPos - 1) / kSfPos;
    const int64_t ntiles = groups * ptiles;
    if (ntiles >= 0x7fffffffll) return fail(BSQ_ERR_ARG, "batch too large for one launch");
    int sms = 0;
    BSQ_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    SfParams q;
    q.bytes = v.bytes;
    q.offs = v.offs;
    q.out = out;
    q.nseq = nseq;
    q.ld = ld;
    q.ntiles = ntiles;
    q.padlen = static_cast<int>(padlen);
    q.ptiles = static_cast<int>(ptiles);
    span_magic(static_cast<uint32_t>(std::max<int64_t>(ptiles, 2)), &q.div_mul, &q.div_shift);  // (unused when ptiles == 1)
    q.ctr = dynamic ? tile_counter_slot(device, st) : nullptr;
    const size_t smem = static_cast<size_t>(kSfStageBytes) + kSfTileBytes;
    const int64_t blocks = std::min<int64_t>(ntiles, static_cast<int64_t>(sms) * std::min(ctas, minb));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(blocks));
    cfg.blockDim = dim3(kSfThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_allowed && q.ctr != nullptr) ? 1 : 0;
    if (minb == 4) BSQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, tokenize_seqfirst_kernel<4>, q, p.lut, p.sp));
    else if (minb == 5) BSQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, tokenize_seqfirst_kernel<5>, q, p.lut, p.sp));
    else BSQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, tokenize_seqfirst_kernel<6>, q, p.lut, p.sp));
    count_launch();
    return BSQ_OK;
}

}  // namespace bsq
