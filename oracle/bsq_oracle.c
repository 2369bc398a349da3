/*
 * bsq_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT part of the product path.
 *
 * Plain-C, single-threaded CPU restatement of the bioseq batch-tokenisation hot path
 * (reference: dnbaker/bioseq, /root/reference/src/alphabet.h + src/tokenize.h +
 * src/tokenize.cpp).  It exists so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py have an independent checker that
 * travels to the GPU box.  Nothing under bioseq_b200/ may import, link or call it.
 *
 * Parity pin: this restatement is checked byte-for-byte against the reference's own
 * compiled tokenizer (oracle/_ref, built by oracle/Makefile from the reference sources
 * where they lie) by tests/test_oracle.py, and against the golden fixtures in
 * tests/golden/ that oracle/make_golden.py generated from that same reference build.
 * The reference repository has no tests of its own for this path; the only published
 * known-answer is README.md:38-44, which is one of the fixtures.
 *
 * Input form: the reference takes a Python sequence of str/bytes/bytearray
 * (src/tokenize.h:389-419); the restatement takes the same residues packed as one
 * byte buffer + int64 offsets (offs[i]..offs[i+1] is sequence i).
 *
 * Defined behaviour where the reference is undefined:
 *   - bytes >= 0x80 index the reference LUT with a negative int8 (alphabet.h:78, UB);
 *     here they are invalid (-1), which is also what the reference build was observed
 *     to produce;
 *   - a sequence with len+bos+eos > padlen aborts the reference process (exception
 *     thrown inside an OpenMP region, tokenize.h:456-459); here it returns an error.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <ctype.h>

#define BSQO_OK 0
#define BSQO_ERR_ARG (-1)
#define BSQO_ERR_TOO_LONG (-2)
#define BSQO_ERR_BAD_TOKEN (-3)
#define BSQO_ERR_KEY (-4)

/* element kinds shared with include/bsq.h */
enum { BSQO_I8 = 0, BSQO_I16 = 1, BSQO_I32 = 2, BSQO_I64 = 3, BSQO_F32 = 4, BSQO_F64 = 5 };

/* ---- alphabet.h:108-124,189-194 (set strings) and :198-222 (CAMAP keys) ---------- */
struct alpha_def { const char *key; const char *set; };
static const struct alpha_def ALPHABETS[] = {
    {"BYTES", NULL},
    {"AMINO20", "A,C,D,E,F,G,H,I,K,L,M,N,P,Q,R,S,T,V,W,Y"},
    {"AMINO", "A,C,D,E,F,G,H,I,K,L,M,N,P,Q,R,S,T,V,W,Y"},
    {"PROTEIN", "A,C,D,E,F,G,H,I,K,L,M,N,P,Q,R,S,T,V,W,Y"},
    {"SEB8", "AST,C,DHN,EKQR,FWY,G,ILMV,P"},
    {"SEB10", "AST,C,DN,EQ,FY,G,HW,ILMV,KR,P"},
    {"SEB14", "A,C,D,EQ,FY,G,H,IV,KR,LM,N,P,ST,W"},
    {"SEV10", "AST,C,DEN,FY,G,H,ILMV,KQR,P,W"},
    {"MURPHY", "A,C,DENQ,FWY,G,H,ILMV,KR,P,ST"},
    {"LIA10", "AC,DE,FWY,G,HN,IV,KQR,LM,P,ST"},
    {"LIB10", "AST,C,DEQ,FWY,G,HN,IV,KR,LM,P"},
    {"SEB6", "AST,CP,DHNEKQR,FWY,G,ILMV"},
    {"DAYHOFF", "AGPST,C,DENQ,FWY,HKR,ILMV"},
    {"DNAMETH", "C,AGT"},
    {"C", "C,AGT"},
    {"KETO", "ACM,KGT"},
    {"PURPYR", "AGR,YCT"},
    {"DNA4", "A,C,G,T"},
    {"DNA", "A,C,G,T"},
    {"DNA5", "A,C,G,T,NMRWSYKVHDB"},
};
#define N_ALPHABETS ((int)(sizeof(ALPHABETS) / sizeof(ALPHABETS[0])))

/* alphabet.h:32-61 make_lut.  Groups are comma separated; group k gets id k; both
 * cases of every member map to the id (:39,:44).  The alias pass (:47-59) is a no-op
 * in the reference: it reads arr[destchar] where destchar is already a token id, so
 * it copies -1 onto -1; it is therefore not restated.  Returns nchars = #groups. */
int bsqo_make_lut(const char *set, int8_t lut[256])
{
    int id = 0;
    memset(lut, 0xff, 256);
    for (const char *p = set; *p; ++p) {
        if (*p == ',') { ++id; continue; }
        unsigned char v = (unsigned char)*p;
        lut[v | 32] = (int8_t)id;
        lut[v & 0xdf] = (int8_t)id;
    }
    return id + 1;
}

/* alphabet.h:92-99 BYTES alphabet: lut[i] = (int8)i, nchars 256.  Values >= 0x80 come
 * out negative and are skipped by the `>= 0` tests downstream. */
static int make_bytes_lut(int8_t lut[256])
{
    for (int i = 0; i < 256; ++i) lut[i] = (int8_t)i;
    return 256;
}

int bsqo_alphabet_count(void) { return N_ALPHABETS; }
const char *bsqo_alphabet_key(int i) { return (i >= 0 && i < N_ALPHABETS) ? ALPHABETS[i].key : NULL; }

/* tokenize.h:72-80: key is upper-cased, then looked up in CAMAP. */
int bsqo_alphabet(const char *key, int8_t lut[256], int *nchars)
{
    char up[32];
    size_t n = strlen(key);
    if (n >= sizeof(up)) return BSQO_ERR_KEY;
    for (size_t i = 0; i <= n; ++i) up[i] = (char)toupper((unsigned char)key[i]);
    for (int i = 0; i < N_ALPHABETS; ++i) {
        if (strcmp(up, ALPHABETS[i].key) == 0) {
            *nchars = ALPHABETS[i].set ? bsqo_make_lut(ALPHABETS[i].set, lut) : make_bytes_lut(lut);
            return BSQO_OK;
        }
    }
    return BSQO_ERR_KEY;
}

/* tokenize.h:22-33 */
void bsqo_ids(int nchars, int eos, int bos, int padchar, int *bos_id, int *eos_id, int *pad_id, int *alphabet_size)
{
    *bos_id = bos ? nchars : -1;
    *eos_id = eos ? nchars + (bos != 0) : -1;
    *pad_id = nchars + (bos != 0) + (eos != 0);
    *alphabet_size = nchars + (bos != 0) + (eos != 0) + (padchar != 0);
}

static void put(void *out, int kind, int64_t idx, int value)
{
    switch (kind) {
    case BSQO_I8:  ((int8_t *)out)[idx] = (int8_t)value; break;
    case BSQO_I16: ((int16_t *)out)[idx] = (int16_t)value; break;
    case BSQO_I32: ((int32_t *)out)[idx] = (int32_t)value; break;
    case BSQO_I64: ((int64_t *)out)[idx] = (int64_t)value; break;
    case BSQO_F32: ((float *)out)[idx] = (float)value; break;
    default:       ((double *)out)[idx] = (double)value; break;
    }
}

static size_t kind_size(int kind)
{
    static const size_t sz[] = {1, 2, 4, 8, 4, 8};
    return (kind >= 0 && kind <= BSQO_F64) ? sz[kind] : 0;
}

/* alphabet.h:78 translate, with bytes >= 0x80 defined as invalid. */
static int translate(const int8_t lut[256], uint8_t c) { return c < 0x80 ? lut[c] : -1; }

/* tokenize.h:381-485 transencode<T>.  bos_id/eos_id < 0 mean "not included";
 * padchar selects whether the tail is filled with pad_id (:473-478) or left at the
 * memset zero (:427).  Invalid residues are skipped (:442) and so stay 0.
 * Output is (n, padlen) if batch_first else (padlen, n)  (:421-425, :430-439). */
int bsqo_tokenize(const uint8_t *bytes, const int64_t *offs, int64_t n, int64_t padlen,
                  const int8_t lut[256], int bos_id, int eos_id, int pad_id, int padchar,
                  int batch_first, int kind, void *out, int64_t *bad_len)
{
    if (padlen <= 0 || kind_size(kind) == 0) return BSQO_ERR_ARG;
    const int bos = bos_id >= 0, eos = eos_id >= 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t tl = offs[i + 1] - offs[i] + bos + eos;
        if (tl > padlen) { if (bad_len) *bad_len = tl; return BSQO_ERR_TOO_LONG; }
    }
    memset(out, 0, (size_t)n * (size_t)padlen * kind_size(kind));
#define AT(s, b) (batch_first ? (b) * padlen + (s) : (s) * n + (b))
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *s = bytes + offs[i];
        const int64_t len = offs[i + 1] - offs[i];
        if (bos) put(out, kind, AT(0, i), bos_id);
        for (int64_t j = 0; j < len; ++j) {
            int tr = translate(lut, s[j]);
            if (tr >= 0) put(out, kind, AT(bos + j, i), tr);
        }
        if (eos) put(out, kind, AT(bos + len, i), eos_id);
        if (padchar)
            for (int64_t k = len + bos + eos; k < padlen; ++k) put(out, kind, AT(k, i), pad_id);
    }
#undef AT
    return BSQO_OK;
}

/* tokenize.h:283-371 tokenize<T>(py::sequence...) = batch_onehot_encode.  Output is
 * always (padlen, n, ncols) (:326); a single 1 per position (:346,:352,:357,:366);
 * invalid or masked-out residues leave the row zero (:349-354); the tail is one-hot
 * pad_id only if padchar (:363-368).  mask (may be NULL) is packed like bytes: one
 * uint8 per residue, 0 = leave the row zero (getmaskptr :372-380). */
int bsqo_onehot(const uint8_t *bytes, const int64_t *offs, const uint8_t *mask, int64_t n,
                int64_t padlen, const int8_t lut[256], int bos_id, int eos_id, int pad_id,
                int padchar, int ncols, int kind, void *out, int64_t *bad_len)
{
    if (padlen <= 0 || kind_size(kind) == 0 || ncols <= 0) return BSQO_ERR_ARG;
    const int bos = bos_id >= 0, eos = eos_id >= 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t tl = offs[i + 1] - offs[i] + bos + eos;
        if (tl > padlen) { if (bad_len) *bad_len = tl; return BSQO_ERR_TOO_LONG; }
    }
    memset(out, 0, (size_t)n * (size_t)padlen * (size_t)ncols * kind_size(kind));
    const int64_t nrc = n * ncols;
#define AT(s, b, c) ((s) * nrc + (b) * ncols + (c))
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *s = bytes + offs[i];
        const uint8_t *m = mask ? mask + offs[i] : NULL;
        const int64_t len = offs[i + 1] - offs[i];
        if (bos) put(out, kind, AT(0, i, bos_id), 1);
        for (int64_t j = 0; j < len; ++j) {
            if (m && !m[j]) continue;
            int tr = translate(lut, s[j]);
            if (tr >= 0) put(out, kind, AT(bos + j, i, tr), 1);
        }
        if (eos) put(out, kind, AT(bos + len, i, eos_id), 1);
        if (padchar)
            for (int64_t k = len + bos + eos; k < padlen; ++k) put(out, kind, AT(k, i, pad_id), 1);
    }
#undef AT
    return BSQO_OK;
}

/* tokenize.h:40-56 / :83-99: id -> first byte value that maps to it (so the upper-case,
 * alphabetically first member of a group); -1 -> the first invalid byte (0x00 for every
 * letter alphabet); specials -> "<BOS>", "<EOS>", "<PAD>".  first[] is indexed by
 * id + 128 to hold the int8 range; returns 1 where an entry exists. */
static void first_bytes(const int8_t lut[256], int16_t first[256])
{
    for (int i = 0; i < 256; ++i) first[i] = -1;
    for (int i = 0; i < 256; ++i) {
        int v = lut[i] + 128;
        if (first[v] < 0) first[v] = (int16_t)i;
    }
}

/* tokenize.h:107-124 load_value then :145/:167 `const uint32_t value = ...` (truncates
 * 8-byte items to their low 32 bits) then lookup.find(value) with an int32 key. */
static int32_t load_key(const uint8_t *p, int itemsize)
{
    switch (itemsize) {
    case 1: return (int32_t)(uint32_t)*p;
    case 2: { uint16_t v; memcpy(&v, p, 2); return (int32_t)(uint32_t)v; }
    case 4: { uint32_t v; memcpy(&v, p, 4); return (int32_t)v; }
    default: { uint64_t v; memcpy(&v, p, 8); return (int32_t)(uint32_t)v; }
    }
}

/* tokenize.h:131-179 decode_tokens for a (rows, cols) array with byte strides.  Writes
 * the concatenated strings to out_chars (capacity out_cap) and rows+1 offsets to
 * out_offs.  A 1-D array is rows = 1.  Returns total chars, BSQO_ERR_BAD_TOKEN with
 * *bad_token = the uint32 value the reference would print (:148,:170), or
 * BSQO_ERR_ARG if out_cap is too small. */
int64_t bsqo_decode(const void *tokens, int itemsize, int64_t rows, int64_t cols,
                    int64_t row_stride, int64_t col_stride, const int8_t lut[256],
                    int bos_id, int eos_id, int pad_id, int padchar,
                    char *out_chars, int64_t out_cap, int64_t *out_offs, uint32_t *bad_token)
{
    int16_t first[256];
    if (itemsize != 1 && itemsize != 2 && itemsize != 4 && itemsize != 8) return BSQO_ERR_ARG;
    first_bytes(lut, first);
    int64_t pos = 0;
    for (int64_t r = 0; r < rows; ++r) {
        out_offs[r] = pos;
        const uint8_t *rp = (const uint8_t *)tokens + r * row_stride;
        for (int64_t c = 0; c < cols; ++c) {
            const int32_t key = load_key(rp + c * col_stride, itemsize);
            const char *sp = NULL;
            /* specials (tokenize.h:91-99) have ids >= nchars, alphabet entries < nchars
             * (BYTES: -128..127), so the two key sets never collide. */
            if (padchar && key == pad_id) sp = "<PAD>";
            if (eos_id >= 0 && key == eos_id) sp = "<EOS>";
            if (bos_id >= 0 && key == bos_id) sp = "<BOS>";
            if (sp) {
                if (pos + 5 > out_cap) return BSQO_ERR_ARG;
                memcpy(out_chars + pos, sp, 5);
                pos += 5;
            } else if (key >= -128 && key <= 127 && first[key + 128] >= 0) {
                if (pos + 1 > out_cap) return BSQO_ERR_ARG;
                out_chars[pos++] = (char)first[key + 128];
            } else {
                if (bad_token) *bad_token = (uint32_t)key;
                return BSQO_ERR_BAD_TOKEN;
            }
        }
    }
    out_offs[rows] = pos;
    return pos;
}

/* ======================================================================================
 * Consumers and augmentation.
 *
 * (B,C,L) one-hot and the embedding gather have no arithmetic of their own: the tests
 * derive their expected values from bsqo_onehot / bsqo_tokenize with a numpy transpose /
 * index, exactly like the reference derives them (bioseq/loaders.py:74-75,
 * bioseq/__init__.py:171-188).
 *
 * BLOSUM62 augmentation: bioseq/blosum.py:36-87.  The reference draws from numpy's PCG64
 * stream (blosum.py:5-6, :61, :80); a GPU cannot reproduce a sequential generator across
 * a batch, so the product path uses the counter-based Philox4x32-10 (Salmon et al., SC'11,
 * "Parallel random numbers: as easy as 1, 2, 3") and this file restates that exact
 * procedure.  Pins: Philox against the Random123 known-answer vectors; the substitution
 * probabilities against the reference's own `normrows` (tests/golden/blosum.json,
 * generated by oracle/make_golden_blosum.py importing bioseq/blosum.py); the procedure's
 * output distribution against those probabilities (tests/test_oracle.py).
 * ====================================================================================== */
void bsqo_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr_in[0], ctr_in[1], ctr_in[2], ctr_in[3]};
    uint32_t k[2] = {key_in[0], key_in[1]};
    for (int round = 0; round < 10; ++round) {
        if (round > 0) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }   /* Weyl key schedule */
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    memcpy(out, c, sizeof(c));
}

/* BLOSUM62 (NCBI), the 21 rows ARNDCQEGHILKMFPSTWYV + X over the 20 amino-acid columns:
 * the block bioseq/blosum.py:36-40 cuts out of the full matrix text (:9-34). */
static const char BLOSUM_ORDER[] = "ARNDCQEGHILKMFPSTWYV";
static const signed char BLOSUM62_ROWS[21][20] = {
    { 4,-1,-2,-2, 0,-1,-1, 0,-2,-1,-1,-1,-1,-2,-1, 1, 0,-3,-2, 0},
    {-1, 5, 0,-2,-3, 1, 0,-2, 0,-3,-2, 2,-1,-3,-2,-1,-1,-3,-2,-3},
    {-2, 0, 6, 1,-3, 0, 0, 0, 1,-3,-3, 0,-2,-3,-2, 1, 0,-4,-2,-3},
    {-2,-2, 1, 6,-3, 0, 2,-1,-1,-3,-4,-1,-3,-3,-1, 0,-1,-4,-3,-3},
    { 0,-3,-3,-3, 9,-3,-4,-3,-3,-1,-1,-3,-1,-2,-3,-1,-1,-2,-2,-1},
    {-1, 1, 0, 0,-3, 5, 2,-2, 0,-3,-2, 1, 0,-3,-1, 0,-1,-2,-1,-2},
    {-1, 0, 0, 2,-4, 2, 5,-2, 0,-3,-3, 1,-2,-3,-1, 0,-1,-3,-2,-2},
    { 0,-2, 0,-1,-3,-2,-2, 6,-2,-4,-4,-2,-3,-3,-2, 0,-2,-2,-3,-3},
    {-2, 0, 1,-1,-3, 0, 0,-2, 8,-3,-3,-1,-2,-1,-2,-1,-2,-2, 2,-3},
    {-1,-3,-3,-3,-1,-3,-3,-4,-3, 4, 2,-3, 1, 0,-3,-2,-1,-3,-1, 3},
    {-1,-2,-3,-4,-1,-2,-3,-4,-3, 2, 4,-2, 2, 0,-3,-2,-1,-2,-1, 1},
    {-1, 2, 0,-1,-3, 1, 1,-2,-1,-3,-2, 5,-1,-3,-1, 0,-1,-3,-2,-2},
    {-1,-1,-2,-3,-1, 0,-2,-3,-2, 1, 2,-1, 5, 0,-2,-1,-1,-1,-1, 1},
    {-2,-3,-3,-3,-2,-3,-3,-3,-1, 0, 0,-3, 0, 6,-4,-2,-2, 1, 3,-1},
    {-1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4, 7,-1,-1,-4,-3,-2},
    { 1,-1, 1, 0,-1, 0, 0, 0,-1,-2,-2, 0,-1,-2,-1, 4, 1,-3,-2,-2},
    { 0,-1, 0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1, 1, 5,-2,-2, 0},
    {-3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1, 1,-4,-3,-2,11, 2,-3},
    {-2,-2,-2,-3,-2,-1,-2,-3, 2,-1,-1,-2,-1, 3,-3,-2,-2, 2, 7,-1},
    { 0,-3,-3,-3,-1,-2,-2,-3,-3, 3, 1,-2, 1,-1,-2,-2, 0,-3,-1, 4},
    { 0,-1,-1,-1,-2,-1,-1,-1,-1,-1,-1,-1,-1,-1,-2, 0, 0,-2,-1,-1},
};

/* normrows of bioseq/blosum.py:41-43: exp2(score) / row sum, as doubles (21 x 20). */
void bsqo_blosum62_probs(double *out)
{
    for (int r = 0; r < 21; ++r) {
        double odds[20], sum = 0.0;
        for (int j = 0; j < 20; ++j) {
            const int s = BLOSUM62_ROWS[r][j];
            odds[j] = s >= 0 ? (double)(1u << s) : 1.0 / (double)(1u << -s);
            sum += odds[j];
        }
        for (int j = 0; j < 20; ++j) out[r * 20 + j] = odds[j] / sum;
    }
}

/* Integer sampling table: thr[r][j] = floor(2^32 * (p[r][0] + ... + p[r][j])), j = 0..18,
 * computed exactly from the integers 16 * 2^score. */
void bsqo_blosum62_thresholds(uint32_t *thr)
{
    for (int r = 0; r < 21; ++r) {
        uint64_t w[20], total = 0, run = 0;
        for (int j = 0; j < 20; ++j) {
            w[j] = (uint64_t)1 << (BLOSUM62_ROWS[r][j] + 4);
            total += w[j];
        }
        for (int j = 0; j < 19; ++j) {
            run += w[j];
            thr[r * 19 + j] = (uint32_t)((run * 4294967296ull) / total);
        }
    }
}

static int blosum_row_of(uint8_t ch)   /* probdict.get(inchar, default_transitions), blosum.py:60 */
{
    for (int j = 0; j < 20; ++j)
        if ((uint8_t)BLOSUM_ORDER[j] == ch) return j;
    return 20;
}

/* augment_seq (blosum.py:63-87) over a packed batch, in place.  Sequence i (global index
 * base + i) is mutated iff augment_frac >= 1 or word 0 of Philox block (g, 0) is below
 * floor(augment_frac * 2^32) (the `rng.uniform() < augment_frac` gate of loaders.py:71).
 * Mutation m retries (blosum.py:79-82) with consecutive (position word, substitution word)
 * pairs taken two per Philox block, blocks numbered from 1 and running on across the
 * chain; position = (word * len) >> 32.  A chain stops retrying after 4096 pairs. */
int bsqo_augment(uint8_t *bytes, const int64_t *offs, int64_t nseq, int chain_len, double augment_frac,
                 uint64_t seed, int64_t base)
{
    uint32_t thr[21 * 19];
    if (chain_len < 0 || !(augment_frac >= 0.0)) return BSQO_ERR_ARG;
    if (chain_len == 0 || augment_frac == 0.0) return BSQO_OK;
    bsqo_blosum62_thresholds(thr);
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const uint64_t gate = augment_frac >= 1.0 ? 4294967296ull : (uint64_t)(augment_frac * 4294967296.0);
    for (int64_t i = 0; i < nseq; ++i) {
        const uint64_t len = (uint64_t)(offs[i + 1] - offs[i]);
        if (len == 0) continue;
        const uint64_t g = (uint64_t)(base + i);
        uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), 0, 0}, rnd[4];
        if (gate < 4294967296ull) {
            bsqo_philox4x32_10(ctr, key, rnd);
            if (rnd[0] >= gate) continue;
        }
        uint8_t *seq = bytes + offs[i];
        uint32_t block = 0;
        for (int m = 0; m < chain_len; ++m) {
            int pairs = 0, mutated = 0;
            while (!mutated && pairs < 4096) {
                ctr[2] = ++block;
                bsqo_philox4x32_10(ctr, key, rnd);
                for (int h = 0; h < 2 && !mutated; ++h, ++pairs) {
                    const uint64_t pos = ((uint64_t)rnd[2 * h] * (len > 0xffffffffull ? 0xffffffffull : len)) >> 32;
                    const uint8_t cur = seq[pos];
                    const uint32_t *t = thr + 19 * blosum_row_of(cur);
                    int j = 0;
                    while (j < 19 && t[j] <= rnd[2 * h + 1]) ++j;
                    if ((uint8_t)BLOSUM_ORDER[j] != cur) {
                        seq[pos] = (uint8_t)BLOSUM_ORDER[j];
                        mutated = 1;
                    }
                }
            }
        }
    }
    return BSQO_OK;
}
