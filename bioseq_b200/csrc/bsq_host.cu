// Host side of libbsq.so: error state, launch counter, the pack layer (ragged host
// sequences -> pinned bytes + int64 offsets) and the host-staged pipeline that overlaps
// host->device copies with the kernels.
//
// Reference counterpart: the serial unpack loop and borrowed-pointer vector of
// src/tokenize.h:386-419 (transencode) / :289-322 (one-hot); the caller-side
// `torch.from_numpy(arr).to(device)` (bioseq/loaders.py:84) is what the staging replaces.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "bsq_internal.h"

namespace bsq {

namespace {
thread_local std::string g_last_error;
thread_local int64_t g_launches = 0;
}  // namespace

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}
void count_launch(int n) { g_launches += n; }

}  // namespace bsq

using bsq::fail;

// ---------------------------------------------------------------------------------------
// pack layer
// ---------------------------------------------------------------------------------------
struct bsq_pack {
    int pinned = 0;
    uint8_t *bytes = nullptr;
    size_t cap_bytes = 0;
    int64_t *offs = nullptr;
    size_t cap_offs = 0;  // entries
    int64_t nseq = 0, nbytes = 0, maxlen = 0;
};

namespace {

int host_alloc(void **p, size_t n, int pinned) {
    if (pinned) {
        BSQ_CUDA_TRY(cudaHostAlloc(p, n, cudaHostAllocPortable));
    } else {
        *p = std::malloc(n);
        if (*p == nullptr) return fail(BSQ_ERR_NOMEM, "out of host memory");
    }
    return BSQ_OK;
}
void host_free(void *p, int pinned) {
    if (p == nullptr) return;
    if (pinned) cudaFreeHost(p);
    else std::free(p);
}

int pack_reserve(bsq_pack *p, int64_t nbytes, int64_t nseq) {
    // +32: the kernels read whole aligned 16-byte words around the residues
    const size_t want_b = static_cast<size_t>(nbytes) + 32, want_o = static_cast<size_t>(nseq) + 1;
    if (want_b > p->cap_bytes) {
        host_free(p->bytes, p->pinned);
        p->bytes = nullptr;
        p->cap_bytes = 0;
        const size_t cap = std::max(want_b + want_b / 4, size_t(1) << 16);
        if (int rc = host_alloc(reinterpret_cast<void **>(&p->bytes), cap, p->pinned)) return rc;
        p->cap_bytes = cap;
    }
    if (want_o > p->cap_offs) {
        host_free(p->offs, p->pinned);
        p->offs = nullptr;
        p->cap_offs = 0;
        const size_t cap = std::max(want_o + want_o / 4, size_t(1) << 10);
        if (int rc = host_alloc(reinterpret_cast<void **>(&p->offs), cap * sizeof(int64_t), p->pinned)) return rc;
        p->cap_offs = cap;
    }
    return BSQ_OK;
}

}  // namespace

extern "C" {

int bsq_pack_create(bsq_pack **out, int pinned) {
    if (out == nullptr) return fail(BSQ_ERR_ARG, "null argument");
    *out = new bsq_pack();
    (*out)->pinned = pinned != 0;
    return BSQ_OK;
}

void bsq_pack_destroy(bsq_pack *p) {
    if (p == nullptr) return;
    host_free(p->bytes, p->pinned);
    host_free(p->offs, p->pinned);
    delete p;
}

int bsq_pack_gather(bsq_pack *p, const void *const *ptrs, const int64_t *lens, int64_t n, int nthreads) {
    if (p == nullptr || n < 0 || (n > 0 && (ptrs == nullptr || lens == nullptr))) return fail(BSQ_ERR_ARG, "bad pack arguments");
    int64_t total = 0, maxlen = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (lens[i] < 0) return fail(BSQ_ERR_ARG, "negative sequence length");
        total += lens[i];
        maxlen = std::max(maxlen, lens[i]);
    }
    if (int rc = pack_reserve(p, total, n)) return rc;
    int64_t acc = 0;
    for (int64_t i = 0; i < n; ++i) {
        p->offs[i] = acc;
        acc += lens[i];
    }
    p->offs[n] = acc;
    p->nseq = n;
    p->nbytes = total;
    p->maxlen = maxlen;

    auto copy_range = [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i)
            if (lens[i] > 0) std::memcpy(p->bytes + p->offs[i], ptrs[i], static_cast<size_t>(lens[i]));
    };
    int nt = std::max(1, nthreads);
    nt = static_cast<int>(std::min<int64_t>(nt, std::max<int64_t>(1, total >> 20)));  // >= 1 MiB per thread
    if (nt <= 1) {
        copy_range(0, n);
    } else {
        // split by bytes, not by count, so ragged batches stay balanced
        std::vector<std::thread> workers;
        int64_t lo = 0;
        for (int t = 0; t < nt; ++t) {
            const int64_t target = total * (t + 1) / nt;
            const int64_t hi = t == nt - 1 ? n : std::upper_bound(p->offs, p->offs + n + 1, target) - p->offs - 1;
            const int64_t hi_c = std::max(lo, std::min(hi, n));
            if (t == nt - 1) copy_range(lo, n);
            else workers.emplace_back(copy_range, lo, hi_c);
            lo = hi_c;
        }
        for (auto &w : workers) w.join();
    }
    std::memset(p->bytes + total, 0, 32);
    return BSQ_OK;
}

const uint8_t *bsq_pack_bytes(const bsq_pack *p) { return p ? p->bytes : nullptr; }
const int64_t *bsq_pack_offsets(const bsq_pack *p) { return p ? p->offs : nullptr; }
int64_t bsq_pack_nseq(const bsq_pack *p) { return p ? p->nseq : 0; }
int64_t bsq_pack_nbytes(const bsq_pack *p) { return p ? p->nbytes : 0; }
int64_t bsq_pack_maxlen(const bsq_pack *p) { return p ? p->maxlen : 0; }

int bsq_check_lengths_host(const int64_t *h_offsets, int64_t nseq, int64_t padlen, const bsq_tokenizer *tok) {
    if (tok == nullptr) return fail(BSQ_ERR_ARG, "null tokenizer");
    if (padlen <= 0) return fail(BSQ_ERR_ARG, "batch tokenize requires padlen is provded.");  // src/tokenize.h:383
    if (nseq < 0 || (nseq > 0 && h_offsets == nullptr)) return fail(BSQ_ERR_ARG, "bad offsets");
    const int64_t extra = (tok->bos_id >= 0) + (tok->eos_id >= 0);
    for (int64_t i = 0; i < nseq; ++i) {
        const int64_t len = h_offsets[i + 1] - h_offsets[i];
        if (len < 0) return fail(BSQ_ERR_ARG, "offsets must be non-decreasing");
        if (len + extra > padlen)  // src/tokenize.h:456-459
            return fail(BSQ_ERR_TOO_LONG, "seq len + bos + eos > padlen: " + std::to_string(len + extra) + ", vs padlen " +
                                              std::to_string(padlen));
    }
    return BSQ_OK;
}

const char *bsq_last_error(void) { return bsq::g_last_error.c_str(); }
int64_t bsq_launch_count(void) { return bsq::g_launches; }
void bsq_launch_count_reset(void) { bsq::g_launches = 0; }

}  // extern "C"

// ---------------------------------------------------------------------------------------
// host-staged pipeline
// ---------------------------------------------------------------------------------------
namespace {
constexpr int kRingSlots = 3;
constexpr size_t kChunkBytes = size_t(4) << 20;  // residues per pipeline stage
constexpr int64_t kSeqAlign = 128;               // chunk boundaries: whole tiles / 16-byte aligned rows
}  // namespace

struct bsq_stager {
    int device = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t done = nullptr;  // completion of the previous call's kernels
    bool busy = false;
    uint8_t *d_bytes = nullptr, *d_mask = nullptr;
    int64_t *d_offs = nullptr;
    size_t cap_bytes = 0, cap_mask = 0, cap_offs = 0;
    uint8_t *ring[kRingSlots] = {nullptr, nullptr, nullptr};
    cudaEvent_t ring_free[kRingSlots] = {nullptr, nullptr, nullptr};
    bool ring_used[kRingSlots] = {false, false, false};
    int ring_next = 0;
    std::vector<cudaEvent_t> events;  // one per chunk in flight
    // BLOSUM62 augmentation applied to every staged range before it is tokenised (chain_len 0 = off)
    int aug_chain = 0;
    double aug_frac = 0.0;
    uint64_t aug_seed = 0;
    int64_t aug_base = 0;
};

namespace {

int dev_reserve(void **p, size_t *cap, size_t want) {
    if (want <= *cap) return BSQ_OK;
    if (*p != nullptr) BSQ_CUDA_TRY(cudaFree(*p));  // implicit device sync: nothing is still reading it
    *p = nullptr;
    *cap = 0;
    const size_t n = want + want / 4 + 256;
    BSQ_CUDA_TRY(cudaMalloc(p, n));
    *cap = n;
    return BSQ_OK;
}

bool is_pinned(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

// host -> device copy of n bytes on the copy stream; pageable sources bounce through the
// pinned ring (the memcpy into slot k overlaps the DMA of slot k-1).
int stage_copy(bsq_stager *s, void *dst, const void *src, size_t n, bool src_pinned) {
    if (n == 0) return BSQ_OK;
    if (src_pinned) {
        BSQ_CUDA_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, s->copy_stream));
        return BSQ_OK;
    }
    size_t done = 0;
    while (done < n) {
        const size_t m = std::min(kChunkBytes, n - done);
        const int k = s->ring_next;
        s->ring_next = (k + 1) % kRingSlots;
        if (s->ring[k] == nullptr) {
            BSQ_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&s->ring[k]), kChunkBytes, cudaHostAllocDefault));
            BSQ_CUDA_TRY(cudaEventCreateWithFlags(&s->ring_free[k], cudaEventDisableTiming));
        }
        if (s->ring_used[k]) BSQ_CUDA_TRY(cudaEventSynchronize(s->ring_free[k]));
        std::memcpy(s->ring[k], static_cast<const uint8_t *>(src) + done, m);
        BSQ_CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(dst) + done, s->ring[k], m, cudaMemcpyHostToDevice, s->copy_stream));
        BSQ_CUDA_TRY(cudaEventRecord(s->ring_free[k], s->copy_stream));
        s->ring_used[k] = true;
        done += m;
    }
    return BSQ_OK;
}

int staged_run(bsq_stager *s, cudaStream_t st, const uint8_t *h_bytes, const int64_t *h_offs, const uint8_t *h_mask,
               int64_t nseq, int64_t padlen, const bsq_tokenizer *tok, int onehot, int batch_first, int kind, void *d_out) {
    if (s == nullptr) return fail(BSQ_ERR_ARG, "null stager");
    if (int rc = bsq::check_launch_args(s->device, nseq, padlen, tok, kind, d_out)) return rc;
    if (int rc = bsq_check_lengths_host(h_offs, nseq, padlen, tok)) {
        // the reference raises invalid_argument for one-hot (:361) and runtime_error for tokens (:458);
        // both map to BSQ_ERR_TOO_LONG here, the Python shim picks the exception type.
        return rc;
    }
    if (nseq == 0) return BSQ_OK;
    const int64_t base = h_offs[0], nbytes = h_offs[nseq] - base;
    if (nbytes > 0 && h_bytes == nullptr) return fail(BSQ_ERR_ARG, "null residue buffer");
    if (s->busy) BSQ_CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, s->done, 0));  // staging buffers still in use
    if (int rc = dev_reserve(reinterpret_cast<void **>(&s->d_bytes), &s->cap_bytes, static_cast<size_t>(nbytes) + 32)) return rc;
    if (int rc = dev_reserve(reinterpret_cast<void **>(&s->d_offs), &s->cap_offs, sizeof(int64_t) * (nseq + 1))) return rc;
    if (h_mask != nullptr)
        if (int rc = dev_reserve(reinterpret_cast<void **>(&s->d_mask), &s->cap_mask, static_cast<size_t>(nbytes) + 32)) return rc;

    const bool pin_b = is_pinned(h_bytes), pin_o = is_pinned(h_offs), pin_m = h_mask && is_pinned(h_mask);
    // offsets first (small); the kernels index d_bytes with (offset - base) via a shifted pointer
    if (int rc = stage_copy(s, s->d_offs, h_offs, sizeof(int64_t) * (nseq + 1), pin_o)) return rc;
    const uint8_t *d_bytes_shifted = s->d_bytes - base;
    const uint8_t *d_mask_shifted = h_mask ? s->d_mask - base : nullptr;
    const size_t esize = bsq_kind_size(kind);
    const int64_t ncols = onehot ? tok->alphabet_size : 1;

    size_t nchunk = 0;
    for (int64_t i0 = 0; i0 < nseq;) {
        // grow the range in whole 128-sequence groups until it holds ~kChunkBytes of residues
        const int64_t want = h_offs[i0] + static_cast<int64_t>(kChunkBytes);
        int64_t i1 = std::upper_bound(h_offs + i0, h_offs + nseq + 1, want) - h_offs - 1;
        i1 = std::max(i1, i0 + 1);
        i1 = std::min(nseq, (i1 + kSeqAlign - 1) / kSeqAlign * kSeqAlign);
        const int64_t b0 = h_offs[i0] - base, b1 = h_offs[i1] - base;
        if (int rc = stage_copy(s, s->d_bytes + b0, h_bytes + base + b0, static_cast<size_t>(b1 - b0), pin_b)) return rc;
        if (h_mask)
            if (int rc = stage_copy(s, s->d_mask + b0, h_mask + base + b0, static_cast<size_t>(b1 - b0), pin_m)) return rc;
        if (nchunk == s->events.size()) {
            cudaEvent_t e;
            BSQ_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s->events.push_back(e);
        }
        BSQ_CUDA_TRY(cudaEventRecord(s->events[nchunk], s->copy_stream));
        BSQ_CUDA_TRY(cudaStreamWaitEvent(st, s->events[nchunk], 0));
        ++nchunk;
        int rc;
        if (s->aug_chain > 0) {
            rc = bsq_augment_blosum62(s->device, st, s->d_bytes - base, s->d_offs + i0, i1 - i0, s->aug_chain, s->aug_frac,
                                      s->aug_seed, s->aug_base + i0);
            if (rc) return rc;
        }
        if (onehot) {
            rc = bsq::launch_onehot(st, d_bytes_shifted, s->d_offs + i0, d_mask_shifted, i1 - i0, nseq, padlen, *tok, kind,
                                    static_cast<uint8_t *>(d_out) + static_cast<size_t>(i0) * ncols * esize);
        } else {
            const size_t off = batch_first ? static_cast<size_t>(i0) * padlen * esize : static_cast<size_t>(i0) * esize;
            rc = bsq::launch_tokenize(st, d_bytes_shifted, s->d_offs + i0, i1 - i0, nseq, padlen, *tok, batch_first, kind,
                                      static_cast<uint8_t *>(d_out) + off, /*first_off=*/h_offs[i0]);
        }
        if (rc) return rc;
        i0 = i1;
    }
    BSQ_CUDA_TRY(cudaEventRecord(s->done, st));
    s->busy = true;
    return BSQ_OK;
}

}  // namespace

extern "C" {

int bsq_stager_create(bsq_stager **out, int device) {
    if (out == nullptr) return fail(BSQ_ERR_ARG, "null argument");
    BSQ_CUDA_TRY(cudaSetDevice(device));
    bsq_stager *s = new bsq_stager();
    s->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete s;
        return fail(BSQ_ERR_CUDA, std::string("bsq_stager_create: ") + cudaGetErrorString(e));
    }
    *out = s;
    return BSQ_OK;
}

void bsq_stager_destroy(bsq_stager *s) {
    if (s == nullptr) return;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    cudaFree(s->d_bytes);
    cudaFree(s->d_mask);
    cudaFree(s->d_offs);
    for (int k = 0; k < kRingSlots; ++k) {
        if (s->ring[k]) cudaFreeHost(s->ring[k]);
        if (s->ring_free[k]) cudaEventDestroy(s->ring_free[k]);
    }
    for (cudaEvent_t e : s->events) cudaEventDestroy(e);
    if (s->done) cudaEventDestroy(s->done);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    delete s;
}

int bsq_stager_sync_copies(bsq_stager *s) {
    if (s == nullptr) return fail(BSQ_ERR_ARG, "null stager");
    BSQ_CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
    return BSQ_OK;
}

int bsq_stager_set_augment(bsq_stager *s, int chain_len, double augment_frac, uint64_t seed, int64_t seq_index_base) {
    if (s == nullptr) return fail(BSQ_ERR_ARG, "null stager");
    if (chain_len < 0 || !(augment_frac >= 0.0)) return fail(BSQ_ERR_ARG, "bad augmentation parameters");
    s->aug_chain = chain_len;
    s->aug_frac = augment_frac;
    s->aug_seed = seed;
    s->aug_base = seq_index_base;
    return BSQ_OK;
}

int bsq_stage_host(bsq_stager *s, void *stream, const uint8_t *h_bytes, const int64_t *h_offsets, int64_t nseq,
                   const uint8_t **d_bytes, const int64_t **d_offsets) {
    if (s == nullptr || d_bytes == nullptr || d_offsets == nullptr) return fail(BSQ_ERR_ARG, "null argument");
    if (nseq < 0 || h_offsets == nullptr) return fail(BSQ_ERR_ARG, "bad offsets");
    BSQ_CUDA_TRY(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t base = h_offsets[0], nbytes = h_offsets[nseq] - base;
    if (nbytes < 0) return fail(BSQ_ERR_ARG, "offsets must be non-decreasing");
    if (nbytes > 0 && h_bytes == nullptr) return fail(BSQ_ERR_ARG, "null residue buffer");
    if (s->busy) BSQ_CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, s->done, 0));
    if (int rc = dev_reserve(reinterpret_cast<void **>(&s->d_bytes), &s->cap_bytes, static_cast<size_t>(nbytes) + 32)) return rc;
    if (int rc = dev_reserve(reinterpret_cast<void **>(&s->d_offs), &s->cap_offs, sizeof(int64_t) * (nseq + 1))) return rc;
    if (int rc = stage_copy(s, s->d_offs, h_offsets, sizeof(int64_t) * (nseq + 1), is_pinned(h_offsets))) return rc;
    if (int rc = stage_copy(s, s->d_bytes, h_bytes + base, static_cast<size_t>(nbytes), nbytes > 0 && is_pinned(h_bytes))) return rc;
    if (s->events.empty()) {
        cudaEvent_t e;
        BSQ_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s->events.push_back(e);
    }
    BSQ_CUDA_TRY(cudaEventRecord(s->events[0], s->copy_stream));
    BSQ_CUDA_TRY(cudaStreamWaitEvent(st, s->events[0], 0));
    if (s->aug_chain > 0 && nseq > 0)
        if (int rc = bsq_augment_blosum62(s->device, st, s->d_bytes - base, s->d_offs, nseq, s->aug_chain, s->aug_frac, s->aug_seed,
                                          s->aug_base))
            return rc;
    // until bsq_stage_release the buffers count as in use by `stream`
    BSQ_CUDA_TRY(cudaEventRecord(s->done, st));
    s->busy = true;
    *d_bytes = s->d_bytes - base;
    *d_offsets = s->d_offs;
    return BSQ_OK;
}

int bsq_stage_release(bsq_stager *s, void *stream) {
    if (s == nullptr) return fail(BSQ_ERR_ARG, "null stager");
    BSQ_CUDA_TRY(cudaSetDevice(s->device));
    BSQ_CUDA_TRY(cudaEventRecord(s->done, static_cast<cudaStream_t>(stream)));
    s->busy = true;
    return BSQ_OK;
}

int bsq_tokenize_host(bsq_stager *s, void *stream, const uint8_t *h_bytes, const int64_t *h_offsets, int64_t nseq,
                      int64_t padlen, const bsq_tokenizer *tok, int batch_first, int kind, void *d_out) {
    return staged_run(s, static_cast<cudaStream_t>(stream), h_bytes, h_offsets, nullptr, nseq, padlen, tok, 0, batch_first,
                      kind, d_out);
}

int bsq_onehot_host(bsq_stager *s, void *stream, const uint8_t *h_bytes, const int64_t *h_offsets, const uint8_t *h_mask,
                    int64_t nseq, int64_t padlen, const bsq_tokenizer *tok, int kind, void *d_out) {
    return staged_run(s, static_cast<cudaStream_t>(stream), h_bytes, h_offsets, h_mask, nseq, padlen, tok, 1, 0, kind, d_out);
}

}  // extern "C"
