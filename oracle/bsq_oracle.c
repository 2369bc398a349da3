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
