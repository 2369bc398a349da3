/*
 * bsq.h -- C ABI of the B200-native batch tokeniser (libbsq.so).
 *
 * This is the drop-in boundary for the hot path of dnbaker/bioseq: everything the
 * reference's pybind11 class `cbioseq.Tokenizer` (src/tokenize.cpp:21-113) computes for
 *     batch_tokenize       (src/tokenize.cpp:82-98  -> Tokenizer::transencode<T>, src/tokenize.h:381-485)
 *     batch_onehot_encode  (src/tokenize.cpp:65-81  -> Tokenizer::tokenize<T>,    src/tokenize.h:283-371)
 *     decode_tokens        (src/tokenize.cpp:49-51  -> Tokenizer::decode_tokens,  src/tokenize.h:131-179)
 * is reachable through the plain-C entry points below.  The reference has no FFI of its
 * own (the Python class *is* its interface); these are the functions a maintainer's
 * pybind11 shim binds instead of the CPU loops, and they are what
 * bioseq_b200's own `cbioseq` module, the GPU parity tests and bench.py call.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch/Python types cross the boundary;
 *   - every function returns 0 (BSQ_OK) or a negative BSQ_ERR_* code and never throws;
 *     bsq_last_error() returns the message of the calling thread's last failure -- where
 *     the reference raises, the text is the reference's text (file:line cited per code);
 *   - "d_" pointers are device memory on `device`, "h_" pointers are host memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *     device entry points only enqueue work, they do not synchronise unless stated;
 *   - sequences travel packed: one byte buffer + int64 offsets, offsets[i]..offsets[i+1]
 *     delimiting sequence i (nseq+1 entries).  This is what the pack layer produces from
 *     the Python str/bytes/bytearray items the reference unpacks at src/tokenize.h:389-419;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails
 *     with BSQ_ERR_CUDA.
 */
#ifndef BSQ_H_
#define BSQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSQ_ABI_VERSION 5

/* ---- status codes ---------------------------------------------------------------- */
#define BSQ_OK 0
#define BSQ_ERR_ARG (-1)       /* std::invalid_argument in the reference -> ValueError          */
#define BSQ_ERR_TOO_LONG (-2)  /* "seq len + bos + eos > padlen: N, vs padlen P" (tokenize.h:458 runtime_error, :361 invalid_argument) */
#define BSQ_ERR_BAD_TOKEN (-3) /* "Unexpected/invalid token N" (tokenize.h:148,170) -> RuntimeError */
#define BSQ_ERR_KEY (-4)       /* "Invalid tokenizer type; select one from..." (tokenize.h:78) -> RuntimeError */
#define BSQ_ERR_CUDA (-5)      /* CUDA runtime failure / no device -> RuntimeError              */
#define BSQ_ERR_NOMEM (-6)
#define BSQ_ERR_IO (-7)        /* file could not be opened / read / written -> RuntimeError     */
#define BSQ_ERR_RANGE (-8)     /* std::out_of_range in the reference -> IndexError              */

/* ---- element kinds of the output array (src/tokenize.cpp:66-79, :83-96) ------------ */
typedef enum bsq_kind {
    BSQ_I8 = 0,  /* destchar b/B : 1-byte integer (reference: int8)            */
    BSQ_I16 = 1, /* destchar h/H */
    BSQ_I32 = 2, /* destchar i/I */
    BSQ_I64 = 3, /* destchar l/L/q/Q : 8-byte integer (reference: uint64)      */
    BSQ_F32 = 4, /* destchar f/F */
    BSQ_F64 = 5  /* destchar d/D */
} bsq_kind;

/* tolower(destchar) dispatch of src/tokenize.cpp:66,83.  Returns a bsq_kind, or
 * BSQ_ERR_ARG ("Unsupported dtype: X", tokenize.cpp:80,97). */
int bsq_kind_of_destchar(char destchar);
/* bytes per element of a kind (1,2,4,8,4,8); 0 for an invalid kind. */
size_t bsq_kind_size(int kind);

/* ---- tokenizer descriptor ---------------------------------------------------------- */
/* Plain data; filled by bsq_tokenizer_init and passed by pointer to the compute calls.
 * Mirrors struct Tokenizer's state (src/tokenize.h:10-38). */
typedef struct bsq_tokenizer {
    int8_t lut[256];        /* alphabet.h:32-61 make_lut: byte -> id, -1 = invalid            */
    int32_t nchars;         /* alphabet.h:27   number of groups                                */
    int32_t bos_id;         /* tokenize.h:23-26  nchars, or -1 when BOS is not included        */
    int32_t eos_id;         /* tokenize.h:27-30  nchars + bos, or -1                           */
    int32_t pad_id;         /* tokenize.h:31-33  nchars + bos + eos (defined even if !padchar) */
    int32_t padchar;        /* tokenize.h:13   zero_onehot_pad_: pad id is a real symbol       */
    int32_t alphabet_size;  /* tokenize.h:22   nchars + eos + bos + padchar                    */
    char key[16];           /* upper-cased key (tokenize.h:73)                                 */
} bsq_tokenizer;

/* Registry of alphabet keys (CAMAP, src/alphabet.h:198-222), in map order. */
int bsq_alphabet_count(void);
const char *bsq_alphabet_key(int index);

/* Tokenizer(key, eos, bos, padchar) -- src/tokenize.h:72-106.  Key lookup is
 * case-insensitive.  BSQ_ERR_KEY on an unknown key. */
int bsq_tokenizer_init(bsq_tokenizer *tok, const char *key, int eos, int bos, int padchar);

/* id -> decoded text, the reference's `lookup` map (src/tokenize.h:83-99): for ids
 * -128..127 the first byte value whose LUT entry is that id; "<BOS>", "<EOS>", "<PAD>"
 * for the specials.  Writes up to cap-1 chars + NUL to buf and returns the length
 * (1 or 5), or 0 when the id has no entry. */
int bsq_tokenizer_lookup(const bsq_tokenizer *tok, int32_t id, char *buf, size_t cap);

/* ---- pack layer (host) -------------------------------------------------------------- */
/* Gathers ragged host sequences into one pinned byte buffer + int64 offsets: the form
 * the kernels read and the form cudaMemcpyAsync can stage without a bounce.  Replaces
 * the reference's vector<pair<const char*, size_t>> of borrowed pointers
 * (src/tokenize.h:386-419). */
typedef struct bsq_pack bsq_pack;

/* pinned != 0: cudaHostAlloc'ed storage (needs a CUDA device); 0: pageable (malloc). */
int bsq_pack_create(bsq_pack **out, int pinned);
void bsq_pack_destroy(bsq_pack *p);
/* Replace the contents with n sequences given as pointer/length arrays; the copy is
 * split over up to nthreads host threads.  Also records the longest length. */
int bsq_pack_gather(bsq_pack *p, const void *const *ptrs, const int64_t *lens, int64_t n, int nthreads);
const uint8_t *bsq_pack_bytes(const bsq_pack *p);
const int64_t *bsq_pack_offsets(const bsq_pack *p); /* n+1 entries */
int64_t bsq_pack_nseq(const bsq_pack *p);
int64_t bsq_pack_nbytes(const bsq_pack *p);
int64_t bsq_pack_maxlen(const bsq_pack *p);

/* Host-side length check done before any launch: the reference aborts the process when
 * len + bos + eos > padlen (exception inside an OpenMP region, src/tokenize.h:456-459);
 * here the call fails with BSQ_ERR_TOO_LONG and the reference's message.
 * padlen <= 0 -> BSQ_ERR_ARG "batch tokenize requires padlen is provded." (tokenize.h:383). */
int bsq_check_lengths_host(const int64_t *h_offsets, int64_t nseq, int64_t padlen, const bsq_tokenizer *tok);
/* Same check for device-resident offsets: one reduction kernel + an 8-byte read-back
 * (synchronises `stream`). */
int bsq_check_lengths_device(int device, void *stream, const int64_t *d_offsets, int64_t nseq,
                             int64_t padlen, const bsq_tokenizer *tok);
/* The same reduction, also validating the offsets themselves before a kernel indexes `d_bytes` with them: every
 * length >= 0, offsets[0] >= 0 and offsets[nseq] <= nbytes (the byte buffer's size; < 0 = unknown, not checked)
 * -> BSQ_ERR_ARG otherwise.  What cbioseq's batch_tokenize_packed / batch_onehot_encode_packed call for
 * device-resident inputs. */
int bsq_check_offsets_device(int device, void *stream, const int64_t *d_offsets, int64_t nseq, int64_t nbytes,
                             int64_t padlen, const bsq_tokenizer *tok);

/* ---- device compute ------------------------------------------------------------------ */
/* batch_tokenize: tokens with BOS/EOS/PAD fused in one pass.
 *   d_out: (nseq, padlen) if batch_first else (padlen, nseq), C-contiguous, `kind`
 *   elements, 16-byte aligned; every element is written exactly once (no memset).
 *   Row i = [bos?] lut[residues] (invalid -> 0) [eos?] then pad_id if padchar else 0
 *   (src/tokenize.h:460-478).  Lengths must already have been checked. */
int bsq_tokenize(int device, void *stream, const uint8_t *d_bytes, const int64_t *d_offsets,
                 int64_t nseq, int64_t padlen, const bsq_tokenizer *tok, int batch_first,
                 int kind, void *d_out);

/* Several batches that share padlen and tokenizer in one launch (batch-first one-byte tokens; other layouts / element
 * types fall back to one launch per batch): batch k = (d_bytes[k], d_offsets[k], nseq[k]) -> d_out[k].  The reference's
 * per-step callers tokenise small batches (training/cnnpretrain.py:123-125,145, bioseq/decoders.py:505): a 4096 x 1000
 * batch is 8 MB of traffic, 1.3 us at the HBM roofline and launch-latency-bound when launched alone; up to 32 of them
 * share one persistent grid here.  The pointer arrays are host arrays and are not referenced after the return.
 * (Single launches are also capturable into a CUDA graph: the kernel then uses its static tile order.) */
int bsq_tokenize_many(int device, void *stream, int nbatch, const uint8_t *const *d_bytes, const int64_t *const *d_offsets,
                      const int64_t *nseq, int64_t padlen, const bsq_tokenizer *tok, int batch_first, int kind,
                      void *const *d_out);

/* batch_onehot_encode: d_out is (padlen, nseq, alphabet_size), always sequence-first
 * (src/tokenize.h:323-326); one 1 per position, all-zero rows for invalid or masked-out
 * residues and -- unless padchar -- for the tail (src/tokenize.h:345-368).
 * d_mask: optional, packed like d_bytes (one uint8 per residue, 0 = zero row). */
int bsq_onehot(int device, void *stream, const uint8_t *d_bytes, const int64_t *d_offsets,
               const uint8_t *d_mask, int64_t nseq, int64_t padlen, const bsq_tokenizer *tok,
               int kind, void *d_out);

/* decode_tokens on a (rows, cols) device array of `itemsize`-byte integers with byte
 * strides (a 1-D array is rows = 1).  Two steps so the caller can size the output:
 *   bsq_decode_lengths: validates every token, fills d_row_offsets[rows+1] (exclusive
 *     prefix sum of the decoded row lengths) and returns the total through *total_chars
 *     (synchronises `stream`).  BSQ_ERR_BAD_TOKEN carries the reference's message with
 *     the first offending value in row-major order (src/tokenize.h:148,170).
 *   bsq_decode_chars: writes the decoded characters of row r to
 *     d_chars[d_row_offsets[r] .. d_row_offsets[r+1]).
 * d_row_tail (optional, `rows` int32 entries of device scratch, the same buffer for both calls): the first step
 * records where each row's trailing run of one repeated special begins (the <PAD>s behind a sequence are 4/5 of
 * the decoded text of a padded batch); the second then writes that run as a pattern fill without reading its
 * tokens again.  NULL for either call: every token is decoded one by one. */
int bsq_decode_lengths(int device, void *stream, const void *d_tokens, int itemsize, int64_t rows,
                       int64_t cols, int64_t row_stride, int64_t col_stride, const bsq_tokenizer *tok,
                       int64_t *d_row_offsets, int32_t *d_row_tail, int64_t *total_chars);
int bsq_decode_chars(int device, void *stream, const void *d_tokens, int itemsize, int64_t rows,
                     int64_t cols, int64_t row_stride, int64_t col_stride, const bsq_tokenizer *tok,
                     const int64_t *d_row_offsets, const int32_t *d_row_tail, uint8_t *d_chars);

/* Both passes in ONE call and one host synchronisation (ABI v5): pass 2 is enqueued right behind pass 1 and reads the
 * total on the device, so the GPU does not idle while the host learns the size and allocates.  The caller provides
 * `capacity` bytes at d_chars (a guess: 3 bytes per token covers batches that are up to half <PAD>; 5 per token
 * always suffices).  Returns like bsq_decode_lengths (*total_chars, d_row_offsets, BSQ_ERR_BAD_TOKEN).  If
 * *total_chars > capacity NOTHING was written to d_chars: allocate *total_chars and call bsq_decode_chars with the
 * offsets and hint this call produced. */
int bsq_decode_text(int device, void *stream, const void *d_tokens, int itemsize, int64_t rows, int64_t cols,
                    int64_t row_stride, int64_t col_stride, const bsq_tokenizer *tok, int64_t *d_row_offsets,
                    int32_t *d_row_tail, uint8_t *d_chars, int64_t capacity, int64_t *total_chars);

/* ---- host-staged entry points (the end-to-end path) ------------------------------------ */
/* A stager owns, for one device: a copy stream, device staging buffers for residues /
 * offsets / mask, and a pinned bounce ring for pageable sources.  The *_host calls split
 * the batch into sequence ranges, cudaMemcpyAsync each range host->device on the copy
 * stream and launch its kernel on `stream` as soon as it lands, so the transfer of
 * range k+1 overlaps the compute of range k.  Inputs may be pinned (copied directly) or
 * pageable (bounced through the ring).  Lengths are checked first (bsq_check_lengths_host).
 * The calls return once all work is enqueued.  Pageable sources have been fully read when
 * the call returns; pinned sources are read asynchronously (cudaMemcpyAsync semantics: do
 * not overwrite them until `stream` reaches this point, or call bsq_stager_sync_copies).
 * The staging buffers stay owned by the stager; two sets alternate call by call, so the
 * copies of a call never wait for the kernels of the call before it. */
typedef struct bsq_stager bsq_stager;
int bsq_stager_create(bsq_stager **out, int device);
void bsq_stager_destroy(bsq_stager *s);
/* Block the host until every host->device copy enqueued by earlier *_host calls on this
 * stager has finished, i.e. until the caller may overwrite the host buffers it passed. */
int bsq_stager_sync_copies(bsq_stager *s);
int bsq_tokenize_host(bsq_stager *s, void *stream, const uint8_t *h_bytes, const int64_t *h_offsets,
                      int64_t nseq, int64_t padlen, const bsq_tokenizer *tok, int batch_first,
                      int kind, void *d_out);
int bsq_onehot_host(bsq_stager *s, void *stream, const uint8_t *h_bytes, const int64_t *h_offsets,
                    const uint8_t *h_mask, int64_t nseq, int64_t padlen, const bsq_tokenizer *tok,
                    int kind, void *d_out);

/* The reference's own calling convention (src/tokenize.h:389-419 / :289-322: a sequence of str / bytes / bytearray
 * items, unpacked into borrowed pointers + lengths) as ONE pass over the items, range by range (~4 MiB of
 * residues each): pool threads walk a range's items, gather exactly the items they walked into a pinned pack
 * and write the offsets as they go; the caller's thread enqueues the range's host->device copies and its kernel
 * as soon as it is gathered.  The DMA of range k overlaps the gather of range k+1 and the walk of range k+2;
 * nothing is walked twice, and back-to-back calls do not synchronise with the copy stream (two pinned packs,
 * owned by the stager, alternate).
 *
 * The items stay with their owner:
 *   resolve(ctx, lo, hi, ptrs, lens)   called from pool threads; fills ptrs[i], lens[i] for lo <= i < hi.  It must
 *                                      only read (no allocation, no interpreter calls); lens[i] = -1 defers item i to
 *   fixup(ctx, i, &ptr, &len)          called on the calling thread (may use the interpreter: the caller holds its
 *                                      lock for the whole call, which is also what keeps the items unchanged while the
 *                                      pool reads them -- the reference holds the GIL throughout as well); non-zero =
 *                                      not an accepted item -> BSQ_ERR_ARG "item was none of string, bytes, ..." (:412).
 * Over-long items -> BSQ_ERR_TOO_LONG with the reference's text (:458), found while streaming: ranges enqueued before
 * the offending one have run, the output is to be discarded.  Returns once all work is enqueued; the items are not
 * referenced after the return. */
typedef void (*bsq_resolve_fn)(void *ctx, int64_t lo, int64_t hi, const void **ptrs, int64_t *lens);
typedef int (*bsq_fixup_fn)(void *ctx, int64_t i, const void **ptr, int64_t *len);
int bsq_tokenize_stream_items(bsq_stager *s, void *stream, int64_t n, bsq_resolve_fn resolve, bsq_fixup_fn fixup, void *ctx,
                              int64_t padlen, const bsq_tokenizer *tok, int batch_first, int kind, void *d_out, int nthreads);
int bsq_onehot_stream_items(bsq_stager *s, void *stream, int64_t n, bsq_resolve_fn resolve, bsq_fixup_fn fixup, void *ctx,
                            int64_t padlen, const bsq_tokenizer *tok, int kind, void *d_out, int nthreads);

/* Same pipeline for items that have already been unpacked into pointer / length arrays.  `p` is not used any more
 * (the stager owns its pinned packs) and may be NULL; ptrs/lens are not referenced after the return. */
int bsq_tokenize_items(bsq_stager *s, bsq_pack *p, void *stream, const void *const *ptrs, const int64_t *lens,
                       int64_t n, int64_t padlen, const bsq_tokenizer *tok, int batch_first, int kind,
                       void *d_out, int nthreads);
int bsq_onehot_items(bsq_stager *s, bsq_pack *p, void *stream, const void *const *ptrs, const int64_t *lens,
                     int64_t n, int64_t padlen, const bsq_tokenizer *tok, int kind, void *d_out, int nthreads);

/* ---- one process, several devices ("one host thread per GPU") ------------------------ */
/* The path shards by sequence index with no exchange step: every output row depends on one input sequence.
 * bsq_shard_bounds splits a packed batch into `nshards` contiguous sequence ranges holding equal shares of the
 * RESIDUES (ragged lengths: equal counts would not balance): bounds[0] = 0 <= ... <= bounds[nshards] = nseq.
 * bsq_tokenize_host_sharded runs shard g (sequences bounds[g] .. bounds[g+1]) through stagers[g] / streams[g]
 * into d_outs[g] -- that device's own (n_g, padlen) / (padlen, n_g) / (padlen, n_g, C) array -- with one pool
 * thread per device enqueuing its copies and kernels concurrently; the host source is shared, nothing is
 * gathered between devices.  onehot != 0 selects batch_onehot_encode (batch_first ignored).  The reference's
 * only multi-GPU construct is nn.DataParallel around the model (training/cnnpretrain.py:86), fed by exactly
 * such per-device batches. */
int bsq_shard_bounds(const int64_t *h_offsets, int64_t nseq, int nshards, int64_t *bounds);
int bsq_tokenize_host_sharded(bsq_stager *const *stagers, void *const *streams, int ndev, const uint8_t *h_bytes,
                              const int64_t *h_offsets, int64_t nseq, int64_t padlen, const bsq_tokenizer *tok, int onehot,
                              int batch_first, int kind, void *const *d_outs, const int64_t *bounds);

/* Run fn(t, nthreads, ctx) for t = 0 .. nthreads-1 on the library's persistent host worker pool (the one the
 * gather / scatter loops above use) and return when all have finished.  The reference's counterpart is its
 * OpenMP team (`#pragma omp parallel for num_threads(nthreads)`, src/tokenize.h:340,452); the Python shim uses
 * this for the per-item pointer/length walk of large batches.  fn must not call back into the pool. */
int bsq_parallel_for(int nthreads, void (*fn)(int t, int nthreads, void *ctx), void *ctx);

/* Second half of decode_tokens (src/tokenize.h:131-179 builds one std::string per row): row r of the
 * decoded characters, d_chars[h_offsets[r] .. h_offsets[r+1]), is delivered to host address dst[r] (the body
 * of the string object the caller created for it).  Device->host copies go through a pinned ring in 8 MiB
 * stages on `stream`; `nthreads` pool threads scatter stage k into the rows while stage k+1 is in flight.
 * h_offsets (rows+1 entries) is on the host.  Blocks until every row has arrived. */
int bsq_fetch_rows(bsq_stager *s, void *stream, const uint8_t *d_chars, const int64_t *h_offsets, int64_t rows,
                   void *const *dst, int nthreads);

/* Stage a packed host batch into the stager's device buffers without running a kernel, for the
 * device entry points that have no *_host twin (bsq_embed, bsq_onehot_bcl, ...).  `stream` is made
 * to wait for the copy.  *d_bytes is biased so that residue k of h_bytes is (*d_bytes)[k], i.e. it
 * pairs with the unrebased *d_offsets exactly like h_bytes pairs with h_offsets.  The buffers stay
 * valid until the next *_host / *_items / bsq_stage_host call on this stager; call bsq_stage_release after
 * enqueuing the kernels that read them so that the next staging waits for those kernels. */
int bsq_stage_host(bsq_stager *s, void *stream, const uint8_t *h_bytes, const int64_t *h_offsets, int64_t nseq,
                   const uint8_t **d_bytes, const int64_t **d_offsets);
int bsq_stage_release(bsq_stager *s, void *stream);

/* ---- consumers fused onto the tokeniser ------------------------------ */
/* One-hot in (nseq, alphabet_size, padlen) layout: what the reference's CNN path builds from
 * batch_onehot_encode with einops.rearrange("length batch emb -> batch emb length") and .float()
 * (bioseq/loaders.py:74-75, :93-94) -- two extra full-tensor passes -- written here in one pass.
 * Element semantics are bsq_onehot's (src/tokenize.h:345-368); d_mask as there. */
int bsq_onehot_bcl(int device, void *stream, const uint8_t *d_bytes, const int64_t *d_offsets, const uint8_t *d_mask,
                   int64_t nseq, int64_t padlen, const bsq_tokenizer *tok, int kind, void *d_out);
/* tokenize -> embedding gather: d_out[(i, p), :] = d_weight[token(i, p), :] with the tokens of
 * bsq_tokenize (never materialised); d_out is (nseq, padlen, row) if batch_first else
 * (padlen, nseq, row).  Replaces batch_tokenize -> torch.from_numpy -> .to(device) -> nn.Embedding
 * (bioseq/__init__.py:171-188 make_embedding and its callers).  Rows are opaque: row_bytes (a
 * multiple of 16) of any element type; nrows >= alphabet_size; d_weight 16-byte aligned. */
int bsq_embed(int device, void *stream, const uint8_t *d_bytes, const int64_t *d_offsets, int64_t nseq, int64_t padlen,
              const bsq_tokenizer *tok, int batch_first, const void *d_weight, int64_t nrows, int64_t row_bytes, void *d_out);

/* ---- BLOSUM62 augmentation on the device ------------------------------ */
/* augment_seq(seq, chain_len) of bioseq/blosum.py:63-87 applied in place to the packed residues of
 * every selected sequence: chain_len times, draw a position and a substitute from the BLOSUM62
 * row of the residue there (probabilities 2^score / row sum, blosum.py:41-43; anything that is not
 * an upper-case amino acid uses the X row, :60) until the substitute differs, and write it.  A
 * sequence is selected with probability augment_frac (always when >= 1), bioseq/loaders.py:71-73.
 * Randomness is Philox4x32-10 keyed by `seed`, counter (seq_index_base + i, block): the result
 * depends only on (seed, global sequence index), not on batching -- and not on numpy's generator,
 * so it matches the reference in distribution, not draw for draw (oracle/bsq_oracle.c restates the
 * exact procedure).  Empty sequences are left alone. */
int bsq_augment_blosum62(int device, void *stream, uint8_t *d_bytes, const int64_t *d_offsets, int64_t nseq, int chain_len,
                         double augment_frac, uint64_t seed, int64_t seq_index_base);
/* The integer substitution table the kernel samples from: thr[21][19] (rows ARNDCQEGHILKMFPSTWYV + X;
 * floor(2^32 * cumulative probability)), row_of[256] (residue byte -> row), aa[20].  NULLs skipped. */
int bsq_blosum62_thresholds(uint32_t *thr, uint8_t *row_of, uint8_t *aa);
/* Apply bsq_augment_blosum62 to every range the *_host / bsq_stage_host calls of this stager stage,
 * between the copy and the kernel (sequence i of a call has index seq_index_base + i).
 * chain_len 0 switches it off. */
int bsq_stager_set_augment(bsq_stager *s, int chain_len, double augment_frac, uint64_t seed, int64_t seq_index_base);

/* ---- FlatFile: the on-disk packed-sequence store that feeds the path -------------------- */
/* The reference's FlatFile (src/fxstats.cpp:26-134) stores a FASTA/FASTQ collection as
 *     uint64 nseqs | uint64 offsets[nseqs+1] | residue bytes          (src/fxstats.cpp:50-59)
 * which is exactly the packed bytes + offsets form the kernels read: a range of sequences
 * [start, stop) of an open file goes to the GPU as
 *     bsq_tokenize_host(stager, stream, bsq_flatfile_bytes(f), bsq_flatfile_offsets(f) + start,
 *                       stop - start, padlen, tok, batch_first, kind, d_out)
 * with no per-sequence host work (the reference materialises one Python bytearray per
 * sequence, src/fxstats.cpp:128-133, and walks them again in src/tokenize.h:389-419). */
typedef struct bsq_flatfile bsq_flatfile;

#define BSQ_FF_MMAP 0   /* map the file read-only (pageable: staged through the pinned ring)        */
#define BSQ_FF_PINNED 1 /* read the file into cudaHostAlloc'ed memory once (direct DMA afterwards) */
#define BSQ_FF_MMAP_PREFAULT 2 /* BSQ_FF_MMAP with the page tables populated at open (MAP_POPULATE): a streamed first
                                  pass over a page-cache-resident file then runs at copy speed instead of fault speed */
#define BSQ_FF_MMAP_REGISTERED 3 /* BSQ_FF_MMAP_PREFAULT, and the mapping is page-locked IN PLACE (cudaHostRegister):
                                    ranges go to the device by direct DMA like BSQ_FF_PINNED, without a second copy of
                                    the file in memory.  Registered read-only where the platform can; else through a
                                    shared writable mapping of the file (never written) when the file may be opened for
                                    writing; else the file stays BSQ_FF_MMAP_PREFAULT (bsq_flatfile_is_pinned tells). */

/* FlatFile::make (src/fxstats.cpp:33-64): parse a FASTA/FASTQ file (plain or gzip; kseq.h
 * record rules) and write the flat file.  outpath NULL or "" -> inpath + ".ff".  Returns the
 * number of sequences and the longest length through the out-pointers (either may be NULL).
 * BSQ_ERR_IO "<path> failed to open" / "<path> could not be opened for writing" (:41,:53),
 * BSQ_ERR_ARG "Cannot handle sequences longer than 2^32 - 1" (:46). */
int bsq_flatfile_make(const char *inpath, const char *outpath, int64_t *nseqs, int64_t *max_seq_len);
/* FlatFile(path, maxseqlen) (src/fxstats.cpp:66-75).  maxseqlen < 0: computed by a scan of
 * the offsets.  Unlike the reference the header is validated against the file size. */
int bsq_flatfile_open(bsq_flatfile **out, const char *path, int64_t maxseqlen, int mode);
void bsq_flatfile_close(bsq_flatfile *f);
int64_t bsq_flatfile_nseqs(const bsq_flatfile *f);
int64_t bsq_flatfile_seq_offset(const bsq_flatfile *f);  /* (nseqs + 2) * 8, src/fxstats.cpp:67 */
int64_t bsq_flatfile_max_seq_len(const bsq_flatfile *f);
const int64_t *bsq_flatfile_offsets(const bsq_flatfile *f); /* nseqs+1 entries, relative to bsq_flatfile_bytes */
const uint8_t *bsq_flatfile_bytes(const bsq_flatfile *f);   /* sequence i = bytes[offsets[i] .. offsets[i+1]) */
int bsq_flatfile_is_pinned(const bsq_flatfile *f);
/* getstats / getlens (src/fxstats.cpp:12-23, :202-219): sequence lengths of a FASTA/FASTQ
 * file.  *lens is malloc'ed by the callee (free it with bsq_free). */
int bsq_fastx_lengths(const char *path, int64_t **lens, int64_t *n);
void bsq_free(void *p);

/* ---- misc -------------------------------------------------------------------------- */
/* Measurement aid: a plain cudaMemcpyAsync device->device on `stream` (bench.py times it on the same rotating
 * buffers as the tokeniser to put a copy of the same HBM traffic next to the kernel's number). */
int bsq_memcpy_d2d(int device, void *stream, void *d_dst, const void *d_src, size_t nbytes);
int bsq_abi_version(void);
const char *bsq_last_error(void);
/* number of kernel launches issued by this library on the calling thread since the last
 * reset (bench.py's gpu_launches counter). */
int64_t bsq_launch_count(void);
void bsq_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* BSQ_H_ */
