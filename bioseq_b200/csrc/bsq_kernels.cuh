// Device-side building blocks of the sm_100a tokeniser kernels.
//
// Everything here is an elementwise byte map bounded by HBM bandwidth; no tensor cores.
// The shared pieces are
//   * a 256-entry byte LUT staged in shared memory (ASCII inputs never bank-conflict:
//     bytes 0..127 live in 32 distinct 4-byte words),
//   * `tokens16`: 16 consecutive output codes of one row -- two 16-byte aligned
//     ld.global.nc loads of the packed residues (clamped to the words that hold the row),
//     an in-register realignment (sequence starts are arbitrary byte offsets and BOS shifts
//     the row by one), 16 LUT look-ups, and BOS / EOS / PAD synthesised by byte masks on the
//     boundary chunks only,
//   * a scalar `token_at` for ragged edges and multi-byte element types.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct bsq_tokenizer;

namespace bsq {

struct LutParam {
    uint32_t w[64];  // 256 byte codes, passed by value in kernel parameter space
};

struct SeqView {
    const uint8_t *bytes;  // packed residues
    const int64_t *offs;   // nseq + 1 offsets into bytes
    const uint8_t *mask;   // optional (one-hot only), indexed like bytes
};

// Byte codes (replicated into all four bytes of a word) for the non-residue columns.
struct Specials {
    uint32_t bos_w, eos_w, pad_w;
    int bos, eos;  // 0/1: symbol present
};

// Tile kernels keep one byte per token in shared memory.  Codes < 0x80 are ids; codes
// 0xFC..0xFF index this 4-entry table (ids that do not fit a byte -- BYTES alphabet --
// and "leave the one-hot row zero").
struct Expand {
    int32_t map[4];  // [0]=pad  [1]=eos  [2]=bos  [3]=invalid
};
constexpr uint32_t kCodePad = 0xFC, kCodeEos = 0xFD, kCodeBos = 0xFE, kCodeInvalid = 0xFF;

// n / d for 0 <= n < 2^31, 1 <= d < 2^31 (Granlund-Montgomery round-up multiplier).
struct FastDiv {
    uint32_t mul, shift, d;
};
__host__ __device__ inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    f.shift = l;
    f.mul = static_cast<uint32_t>((((1ull << l) - d) << 32) / d + 1);
    return f;
}
__host__ __device__ __forceinline__ uint32_t fd_div(uint32_t n, const FastDiv &f) {
    const uint32_t t = static_cast<uint32_t>((static_cast<uint64_t>(n) * f.mul) >> 32);
    return (t + n) >> f.shift;
}

// Kernel-side form of a tokenizer (built on the host by prepare(), bsq_kernels.cu):
// mode 0: batch-first one-byte tokens (codes are the output bytes), mode 1: tokens through Expand
// (wide element types, tile kernels, embedding gather), mode 2: one-hot columns through Expand.
struct Prepared {
    LutParam lut;
    Specials sp;
    Expand ex;
};
Prepared prepare(const bsq_tokenizer &tok, int mode);

// K1s (bsq_span.cu): batch-first one-byte tokens, tile-staged and warp-specialised.
bool span_kernel_applicable(int64_t padlen);
int sm_count(int device);  // SMs of the device (cached per device)
inline int cur_sms() {     // ... of the current device (the entry points have made the call's device current)
    int d = 0;
    cudaGetDevice(&d);
    return sm_count(d);
}
int launch_tokenize_span(int device, cudaStream_t st, const SeqView &v, int64_t nseq, int64_t padlen, const Prepared &p,
                         uint8_t *out, bool pdl_allowed);
// A slot of the per-stream tile counters of the dynamic schedulers (bsq_span.cu), or nullptr when none can be used
// (stream capture in progress, out of memory): the kernels then keep their static tile order.
unsigned int *tile_counter_slot(int device, cudaStream_t st);
// n / d = umulhi(n, mul) >> shift for 0 <= n < 2^31, 2 <= d <= 2^30.
void span_magic(uint32_t d, uint32_t *mul, uint32_t *shift);
constexpr int kSpanManyMax = 32;  // batches per launch of launch_tokenize_span_many
int launch_tokenize_span_many(int device, cudaStream_t st, int nbatch, const uint8_t *const *d_bytes, const int64_t *const *d_offs,
                              const int64_t *nseqs, int64_t padlen, const Prepared &p, uint8_t *const *outs, bool pdl_allowed);

#ifdef __CUDACC__

__device__ __forceinline__ uint4 ldg16(uintptr_t addr) {
    return __ldg(reinterpret_cast<const uint4 *>(addr));
}

__device__ __forceinline__ void load_lut(uint8_t *lut_smem, const LutParam &p) {
    if (threadIdx.x < 64) reinterpret_cast<uint32_t *>(lut_smem)[threadIdx.x] = p.w[threadIdx.x];
}

// Mask of the bytes j < k of a 16-byte chunk that fall in 32-bit word `w` (j = 4w..4w+3):
// the high word of (0:ffffffff) << clamp(8*(k-4w), 0, 32) -- one VIADDMNMX + one SHF.L.W.
// k must already be clamped to [0, 16].
__device__ __forceinline__ uint32_t lt_mask(int k, int w) {
    return __funnelshift_lc(0xffffffffu, 0u, static_cast<uint32_t>(max(8 * k - 32 * w, 0)));
}
__device__ __forceinline__ int clamp16(int k) { return min(max(k, 0), 16); }

// Four LUT look-ups: byte extraction with PRMT (one ALU op each), LDS.U8, and the re-packing as
// integer multiply-adds so it lands on the FMA pipe instead of the (half-rate) ALU pipe.
__device__ __forceinline__ uint32_t translate4(uint32_t x, const uint8_t *lut) {
    const uint32_t b0 = lut[__byte_perm(x, 0u, 0x4440)];
    const uint32_t b1 = lut[__byte_perm(x, 0u, 0x4441)];
    const uint32_t b2 = lut[__byte_perm(x, 0u, 0x4442)];
    const uint32_t b3 = lut[__byte_perm(x, 0u, 0x4443)];
    return b0 + b1 * 0x100u + b2 * 0x10000u + b3 * 0x1000000u;
}

// Where the bytes of one row live: column c of the row (BOS shift included) is al[off + c].
// `al` is 16-byte aligned; fw/lw are the byte offsets (multiples of 16, relative to al) of the
// first and last aligned 16-byte words that hold a residue of this row.  Loads are clamped to
// [fw, lw], so nothing outside words containing valid residues is ever touched, whatever
// column range is asked for (a clamped word only ever replaces bytes that are masked out).
struct RowSrc {
    const uint8_t *al;
    int off, fw, lw;
};
__device__ __forceinline__ RowSrc make_rowsrc(const uint8_t *base, int64_t start, int bos, int len) {
    const uint8_t *src = base + start - bos;
    RowSrc r;
    r.off = static_cast<int>(reinterpret_cast<uintptr_t>(src) & 15u);
    r.al = src - r.off;
    r.fw = (r.off + bos) & ~15;
    r.lw = (r.off + bos + len - 1) & ~15;
    return r;
}

// The 16 bytes of columns c0 .. c0+15, in two steps so that callers can issue the loads of
// several chunks before consuming any of them:
//   fetch_issue   two aligned 16-byte ld.global.nc loads (the second only when the chunk is not
//                 16-byte aligned in the source),
//   fetch_align   funnel-shift realignment into four little-endian words.
// The byte shift (off + c0) & 15 is the same for every chunk of a row, so when a warp works on
// one row the switch is warp-uniform.
struct Fetched {
    uint4 v0, v1;
};
__device__ __forceinline__ Fetched fetch_issue(const RowSrc &rs, int c0) {
    const int a = rs.off + c0;
    const int a0 = a & ~15;
    Fetched f;
    f.v0 = ldg16(reinterpret_cast<uintptr_t>(rs.al) + static_cast<uint32_t>(min(max(a0, rs.fw), rs.lw)));
    f.v1 = make_uint4(0u, 0u, 0u, 0u);
    if ((a & 15) != 0)
        f.v1 = ldg16(reinterpret_cast<uintptr_t>(rs.al) + static_cast<uint32_t>(min(max(a0 + 16, rs.fw), rs.lw)));
    return f;
}
__device__ __forceinline__ void fetch_align(const Fetched &f, const RowSrc &rs, int c0, uint32_t out[4]) {
    const uint32_t s = static_cast<uint32_t>(rs.off + c0) & 15u;
    const uint32_t sh = (s & 3u) * 8u;
    const uint4 &v0 = f.v0, &v1 = f.v1;
    switch (s >> 2) {
        case 0:
            out[0] = __funnelshift_r(v0.x, v0.y, sh); out[1] = __funnelshift_r(v0.y, v0.z, sh);
            out[2] = __funnelshift_r(v0.z, v0.w, sh); out[3] = __funnelshift_r(v0.w, v1.x, sh);
            break;
        case 1:
            out[0] = __funnelshift_r(v0.y, v0.z, sh); out[1] = __funnelshift_r(v0.z, v0.w, sh);
            out[2] = __funnelshift_r(v0.w, v1.x, sh); out[3] = __funnelshift_r(v1.x, v1.y, sh);
            break;
        case 2:
            out[0] = __funnelshift_r(v0.z, v0.w, sh); out[1] = __funnelshift_r(v0.w, v1.x, sh);
            out[2] = __funnelshift_r(v1.x, v1.y, sh); out[3] = __funnelshift_r(v1.y, v1.z, sh);
            break;
        default:
            out[0] = __funnelshift_r(v0.w, v1.x, sh); out[1] = __funnelshift_r(v1.x, v1.y, sh);
            out[2] = __funnelshift_r(v1.y, v1.z, sh); out[3] = __funnelshift_r(v1.z, v1.w, sh);
            break;
    }
}
__device__ __forceinline__ void fetch16(const RowSrc &rs, int c0, uint32_t out[4]) {
    const Fetched f = fetch_issue(rs, c0);
    fetch_align(f, rs, c0, out);
}
// does the chunk starting at column c0 hold at least one residue?
__device__ __forceinline__ bool has_residues(int c0, int bos, int len) {
    const int r0 = c0 - bos;
    return min(r0 + 16, len) > max(r0, 0);
}

// Per-CTA shared-memory tables for the chunk that holds the end of a row: for k = 0..16 leading
// bytes that are still BOS/residues, m[k] keeps those bytes and f[k] supplies the rest (EOS at
// byte k when the tokenizer has one, the pad code after it).  Two LDS.128 + four LOP3 replace
// ~70 mask-building instructions per tail chunk.
struct TailTab {
    uint4 m[17];
    uint4 f[17];
};
__device__ __forceinline__ void init_tailtab(TailTab &tab, const Specials &sp) {
    const int k = static_cast<int>(threadIdx.x) - 64;  // threads 64..80 (0..63 load the LUT)
    if (k >= 0 && k <= 16) {
        uint32_t m[4], f[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            m[w] = lt_mask(k, w);
            const uint32_t m_eos = lt_mask(clamp16(k + sp.eos), w);
            f[w] = (sp.eos_w & m_eos & ~m[w]) | (sp.pad_w & ~m_eos);
        }
        tab.m[k] = make_uint4(m[0], m[1], m[2], m[3]);
        tab.f[k] = make_uint4(f[0], f[1], f[2], f[3]);
    }
}

// Codes of columns c0 .. c0+15 of a row with `len` residues, for c0 < bos + len + eos (the
// caller emits the constant pad vector beyond that).  Column layout (src/tokenize.h:460-478):
//   [0, bos)            BOS
//   [bos, bos+len)      lut[residue]
//   bos+len             EOS (if eos)
//   beyond              pad code
// With MAYBE_NEG, c0 may be negative (rows that do not start 16-byte aligned in the output);
// bytes left of column 0 are don't-care.  `ms` is the row's mask source (one-hot only) or nullptr.
template <bool MAYBE_NEG>
__device__ __forceinline__ uint4 tokens16_finish(uint32_t t[4], int len, int c0, const Specials &sp, const TailTab &tab) {
    const int n = sp.bos + len;  // column of EOS
    if (MAYBE_NEG && c0 < 0) {
        const int kr0 = clamp16(sp.bos - c0), kr1 = clamp16(n - c0), kr2 = clamp16(n + sp.eos - c0);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t m_bos = lt_mask(kr0, w);
            const uint32_t m_res = lt_mask(kr1, w);
            const uint32_t m_eos = lt_mask(kr2, w);
            t[w] = (sp.bos_w & m_bos) | (t[w] & m_res & ~m_bos) | (sp.eos_w & m_eos & ~m_res) | (sp.pad_w & ~m_eos);
        }
    } else {
        if (c0 < sp.bos) t[0] = __byte_perm(t[0], sp.bos_w, 0x3214);  // c0 == 0: column 0 is BOS
        if (c0 + 16 > n) {                                           // the row ends inside this chunk
            const uint4 m = tab.m[n - c0], f = tab.f[n - c0];
            t[0] = (t[0] & m.x) | f.x; t[1] = (t[1] & m.y) | f.y;
            t[2] = (t[2] & m.z) | f.z; t[3] = (t[3] & m.w) | f.w;
        }
    }
    return make_uint4(t[0], t[1], t[2], t[3]);
}

template <bool MAYBE_NEG>
__device__ __forceinline__ uint4 tokens16(const RowSrc &rs, const RowSrc *ms, int len, int c0,
                                          const Specials &sp, const uint8_t *lut, const TailTab &tab) {
    uint32_t t[4] = {0u, 0u, 0u, 0u};
    if (has_residues(c0, sp.bos, len)) {
        uint32_t raw[4];
        fetch16(rs, c0, raw);
#pragma unroll
        for (int w = 0; w < 4; ++w) t[w] = translate4(raw[w], lut);
        if (ms != nullptr) {  // masked-out residues become kCodeInvalid (0xFF)
            uint32_t m[4];
            fetch16(*ms, c0, m);
#pragma unroll
            for (int w = 0; w < 4; ++w) t[w] |= __vcmpeq4(m[w], 0u);
        }
    }
    return tokens16_finish<MAYBE_NEG>(t, len, c0, sp, tab);
}

// Same result as tokens16 (no mask) from loads issued earlier with fetch_issue.
template <bool MAYBE_NEG>
__device__ __forceinline__ uint4 tokens16_from(const Fetched &f, bool loaded, const RowSrc &rs, int len, int c0,
                                               const Specials &sp, const uint8_t *lut, const TailTab &tab) {
    uint32_t t[4] = {0u, 0u, 0u, 0u};
    if (loaded) {
        uint32_t raw[4];
        fetch_align(f, rs, c0, raw);
#pragma unroll
        for (int w = 0; w < 4; ++w) t[w] = translate4(raw[w], lut);
    }
    return tokens16_finish<MAYBE_NEG>(t, len, c0, sp, tab);
}

// One code, any column: the scalar twin of tokens16.
__device__ __forceinline__ uint32_t token_at(const SeqView &v, int64_t start, int len, int c,
                                             const Specials &sp, const uint8_t *lut) {
    const int r = c - sp.bos;
    if (r < 0) return sp.bos_w & 0xffu;
    if (r < len) {
        uint32_t code = lut[__ldg(v.bytes + start + r)];
        if (v.mask != nullptr && __ldg(v.mask + start + r) == 0) code = kCodeInvalid;
        return code;
    }
    if (r == len && sp.eos) return sp.eos_w & 0xffu;
    return sp.pad_w & 0xffu;
}

__device__ __forceinline__ int32_t expand_code(uint32_t code, const Expand &ex) {
    return code < 0x80u ? static_cast<int32_t>(code) : ex.map[code & 3u];
}

#endif  // __CUDACC__

}  // namespace bsq
