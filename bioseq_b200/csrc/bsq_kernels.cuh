// Device-side building blocks of the sm_100a tokeniser kernels.
//
// Everything here is an elementwise byte map bounded by HBM bandwidth; no tensor cores.
// The shared pieces are
//   * a 256-entry byte LUT staged in shared memory (ASCII inputs never bank-conflict:
//     bytes 0..127 live in 32 distinct 4-byte words),
//   * `tokens16`: 16 consecutive output codes of one row -- two 16-byte aligned
//     ld.global.nc loads of the packed residues, an in-register realignment (sequence
//     starts are arbitrary byte offsets and BOS shifts the row by one), 16 LUT look-ups,
//     and BOS / EOS / PAD synthesised by byte masks on the boundary chunks only,
//   * a scalar `token_at` for ragged edges and multi-byte element types.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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

#ifdef __CUDACC__

__device__ __forceinline__ uint4 ldg16(uintptr_t addr) {
    return __ldg(reinterpret_cast<const uint4 *>(addr));
}

__device__ __forceinline__ void load_lut(uint8_t *lut_smem, const LutParam &p) {
    if (threadIdx.x < 64) reinterpret_cast<uint32_t *>(lut_smem)[threadIdx.x] = p.w[threadIdx.x];
}

// Mask of the bytes j < k of a 16-byte chunk that fall in 32-bit word `w` (j = 4w..4w+3).
__device__ __forceinline__ uint32_t lt_mask(int k, int w) {
    const int kk = min(max(k - 4 * w, 0), 4);
    return kk >= 4 ? 0xffffffffu : ((1u << (8 * kk)) - 1u);
}

__device__ __forceinline__ uint32_t translate4(uint32_t x, const uint8_t *lut) {
    const uint32_t b0 = lut[x & 0xffu];
    const uint32_t b1 = lut[(x >> 8) & 0xffu];
    const uint32_t b2 = lut[(x >> 16) & 0xffu];
    const uint32_t b3 = lut[x >> 24];
    return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}

// The 16 bytes base[first .. first+16) as four little-endian words, fetched with at most
// two aligned 16-byte loads.  Only aligned words that overlap the needed index range
// [lo, hi) are touched, so nothing outside the pages holding valid bytes is ever read
// (first may be lo-1 when a BOS column precedes the residues).
__device__ __forceinline__ void fetch16(const uint8_t *base, int64_t first, int64_t lo, int64_t hi, uint32_t out[4]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(base) + first;
    const uintptr_t a0 = a & ~static_cast<uintptr_t>(15);
    const uintptr_t need_lo = reinterpret_cast<uintptr_t>(base) + lo;
    const uintptr_t need_hi = reinterpret_cast<uintptr_t>(base) + hi;
    uint4 v0 = make_uint4(0, 0, 0, 0), v1 = make_uint4(0, 0, 0, 0);
    if (a0 + 16 > need_lo) v0 = ldg16(a0);
    if (a0 + 16 < need_hi) v1 = ldg16(a0 + 16);
    const uint32_t s = static_cast<uint32_t>(a) & 15u;
    uint32_t x0 = v0.x, x1 = v0.y, x2 = v0.z, x3 = v0.w, x4 = v1.x, x5 = v1.y;
    if (s & 8u) { x0 = x2; x1 = x3; x2 = v1.x; x3 = v1.y; x4 = v1.z; x5 = v1.w; }
    if (s & 4u) { x0 = x1; x1 = x2; x2 = x3; x3 = x4; x4 = x5; }
    const uint32_t sh = (s & 3u) * 8u;
    out[0] = __funnelshift_r(x0, x1, sh);
    out[1] = __funnelshift_r(x1, x2, sh);
    out[2] = __funnelshift_r(x2, x3, sh);
    out[3] = __funnelshift_r(x3, x4, sh);
}

// Codes of columns c0 .. c0+15 (c0 >= 0) of the row whose residues are
// bytes[start .. start+len).  Column layout (src/tokenize.h:460-478):
//   [0, bos)            BOS
//   [bos, bos+len)      lut[residue]
//   bos+len             EOS (if eos)
//   beyond              pad code
// Columns past the row's padlen come out as pad and are ignored by the callers.
__device__ __forceinline__ uint4 tokens16(const SeqView &v, int64_t start, int len, int c0,
                                          const Specials &sp, const uint8_t *lut) {
    const int r0 = c0 - sp.bos;  // residue index of the chunk's first column (may be -1)
    const int lo = max(r0, 0), hi = min(r0 + 16, len);
    uint32_t t[4] = {0u, 0u, 0u, 0u};
    if (hi > lo) {
        uint32_t raw[4];
        fetch16(v.bytes, start + r0, start + lo, start + hi, raw);
#pragma unroll
        for (int w = 0; w < 4; ++w) t[w] = translate4(raw[w], lut);
        if (v.mask != nullptr) {  // one-hot only: masked-out residues become kCodeInvalid
            uint32_t m[4];
            fetch16(v.mask, start + r0, start + lo, start + hi, m);
#pragma unroll
            for (int w = 0; w < 4; ++w) t[w] |= __vcmpeq4(m[w], 0u);
        }
    }
    const int n = sp.bos + len;  // column of EOS
    if (c0 < sp.bos || c0 + 16 > n) {
        const int kr0 = sp.bos - c0, kr1 = n - c0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t m_bos = lt_mask(kr0, w);
            const uint32_t m_res = lt_mask(kr1, w);
            const uint32_t m_eos = lt_mask(kr1 + sp.eos, w);
            t[w] = (sp.bos_w & m_bos) | (t[w] & m_res & ~m_bos) | (sp.eos_w & m_eos & ~m_res) | (sp.pad_w & ~m_eos);
        }
    }
    return make_uint4(t[0], t[1], t[2], t[3]);
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
