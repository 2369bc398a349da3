// Alphabet registry, LUT construction and tokenizer ids (host side of libbsq.so).
//
// Reference behaviour mirrored here:
//   src/alphabet.h:32-61   TAlphabet::make_lut -- comma separated groups, group k -> id k,
//                          both letter cases; the alias pass (":47-59") never changes the
//                          table in the reference (it indexes the LUT with a token id), so
//                          'U' stays invalid in DNA and 'O'/'U' in the protein alphabets.
//   src/alphabet.h:92-99   BYTES: lut[i] = int8(i), 256 symbols.
//   src/alphabet.h:198-222 CAMAP: the 20 registered keys.
//   src/tokenize.h:22-38   special-symbol ids, :72-106 constructor, :83-99 lookup map.
#include "bsq_internal.h"

#include <algorithm>
#include <cctype>
#include <cstring>
#include <string>

namespace {

struct AlphabetDef {
    const char *key;     // registry key (upper case)
    const char *groups;  // comma separated residue groups; nullptr = identity (BYTES)
};

constexpr const char *kAmino20 = "A,C,D,E,F,G,H,I,K,L,M,N,P,Q,R,S,T,V,W,Y";
constexpr const char *kMeth = "C,AGT";
constexpr const char *kDna4 = "A,C,G,T";

// std::map iteration order of CAMAP (lexicographic) -- also the order the reference lists
// the options in its "Invalid tokenizer type" message.
constexpr AlphabetDef kRegistry[] = {
    {"AMINO", kAmino20},
    {"AMINO20", kAmino20},
    {"BYTES", nullptr},
    {"C", kMeth},
    {"DAYHOFF", "AGPST,C,DENQ,FWY,HKR,ILMV"},
    {"DNA", kDna4},
    {"DNA4", kDna4},
    {"DNA5", "A,C,G,T,NMRWSYKVHDB"},
    {"DNAMETH", kMeth},
    {"KETO", "ACM,KGT"},
    {"LIA10", "AC,DE,FWY,G,HN,IV,KQR,LM,P,ST"},
    {"LIB10", "AST,C,DEQ,FWY,G,HN,IV,KR,LM,P"},
    {"MURPHY", "A,C,DENQ,FWY,G,H,ILMV,KR,P,ST"},
    {"PROTEIN", kAmino20},
    {"PURPYR", "AGR,YCT"},
    {"SEB10", "AST,C,DN,EQ,FY,G,HW,ILMV,KR,P"},
    {"SEB14", "A,C,D,EQ,FY,G,H,IV,KR,LM,N,P,ST,W"},
    {"SEB6", "AST,CP,DHNEKQR,FWY,G,ILMV"},
    {"SEB8", "AST,C,DHN,EKQR,FWY,G,ILMV,P"},
    {"SEV10", "AST,C,DEN,FY,G,H,ILMV,KQR,P,W"},
};
constexpr int kNumAlphabets = int(sizeof(kRegistry) / sizeof(kRegistry[0]));

int build_lut(const AlphabetDef &def, int8_t lut[256]) {
    if (def.groups == nullptr) {
        for (int b = 0; b < 256; ++b) lut[b] = static_cast<int8_t>(b);
        return 256;
    }
    std::fill(lut, lut + 256, int8_t(-1));
    int group = 0;
    for (const char *p = def.groups; *p != '\0'; ++p) {
        const unsigned char ch = static_cast<unsigned char>(*p);
        if (ch == ',') {
            ++group;
        } else {
            lut[ch | 0x20] = static_cast<int8_t>(group);   // lower case
            lut[ch & 0xdf] = static_cast<int8_t>(group);   // upper case
        }
    }
    return group + 1;
}

}  // namespace

extern "C" {

int bsq_alphabet_count(void) { return kNumAlphabets; }

const char *bsq_alphabet_key(int index) {
    return (index >= 0 && index < kNumAlphabets) ? kRegistry[index].key : nullptr;
}

int bsq_tokenizer_init(bsq_tokenizer *tok, const char *key, int eos, int bos, int padchar) {
    if (tok == nullptr || key == nullptr) return bsq::fail(BSQ_ERR_ARG, "bsq_tokenizer_init: null argument");
    std::string up(key);
    std::transform(up.begin(), up.end(), up.begin(), [](unsigned char c) { return char(std::toupper(c)); });
    const AlphabetDef *def = nullptr;
    for (const AlphabetDef &d : kRegistry)
        if (up == d.key) def = &d;
    if (def == nullptr) {
        std::string msg = "Invalid tokenizer type; select one from";   // src/tokenize.h:76-78
        for (const AlphabetDef &d : kRegistry) { msg += d.key; msg += ';'; }
        return bsq::fail(BSQ_ERR_KEY, msg);
    }
    std::memset(tok, 0, sizeof(*tok));
    tok->nchars = build_lut(*def, tok->lut);
    const int b = bos != 0, e = eos != 0, p = padchar != 0;
    tok->bos_id = b ? tok->nchars : -1;
    tok->eos_id = e ? tok->nchars + b : -1;
    tok->pad_id = tok->nchars + b + e;
    tok->padchar = p;
    tok->alphabet_size = tok->nchars + b + e + p;
    std::strncpy(tok->key, def->key, sizeof(tok->key) - 1);
    return BSQ_OK;
}

int bsq_tokenizer_lookup(const bsq_tokenizer *tok, int32_t id, char *buf, size_t cap) {
    if (tok == nullptr || buf == nullptr || cap < 6) return 0;
    const char *special = nullptr;
    if (tok->bos_id >= 0 && id == tok->bos_id) special = "<BOS>";
    else if (tok->eos_id >= 0 && id == tok->eos_id) special = "<EOS>";
    else if (tok->padchar && id == tok->pad_id) special = "<PAD>";
    if (special != nullptr) {
        std::memcpy(buf, special, 6);
        return 5;
    }
    if (id < -128 || id > 127) return 0;
    for (int b = 0; b < 256; ++b) {
        if (tok->lut[b] == id) {
            buf[0] = static_cast<char>(b);
            buf[1] = '\0';
            return 1;
        }
    }
    return 0;
}

int bsq_kind_of_destchar(char destchar) {
    switch (std::tolower(static_cast<unsigned char>(destchar))) {   // src/tokenize.cpp:66,83
        case 'b': return BSQ_I8;
        case 'h': return BSQ_I16;
        case 'i': return BSQ_I32;
        case 'l':
        case 'q': return BSQ_I64;
        case 'f': return BSQ_F32;
        case 'd': return BSQ_F64;
        default: break;
    }
    return bsq::fail(BSQ_ERR_ARG, std::string("Unsupported dtype: ") + destchar);
}

size_t bsq_kind_size(int kind) {
    static const size_t sizes[] = {1, 2, 4, 8, 4, 8};
    return (kind >= 0 && kind <= BSQ_F64) ? sizes[kind] : 0;
}

int bsq_abi_version(void) { return BSQ_ABI_VERSION; }

}  // extern "C"
