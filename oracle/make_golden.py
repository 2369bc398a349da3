"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/golden.json from the reference's own
compiled tokenizer (oracle/_ref/ref_cbioseq, see oracle/Makefile).  Run in the build
container (needs /root/reference to have been compiled):

    make -C oracle ref && python oracle/make_golden.py

The fixture holds (i) the signed 256-entry LUT, nchars and special ids of every registered
alphabet key, (ii) literal small known-answer cases (README.md:38-44 and the SURVEY 8c
probes), (iii) SHA-256 digests + shapes of reference outputs on seeded synthetic batches
(generator: bioseq_b200/synth.py).  The committed JSON is what travels to the GPU box.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bioseq_b200.synth import gen, gen_mask, as_list, AA20  # noqa: E402
from oracle.oracle import load_ref  # noqa: E402

R = load_ref()
assert R is not None, "build the reference first: make -C oracle ref"

KEYS = ["BYTES", "AMINO20", "AMINO", "PROTEIN", "SEB8", "SEB10", "SEB14", "SEV10", "MURPHY", "LIA10",
        "LIB10", "SEB6", "DAYHOFF", "DNAMETH", "C", "KETO", "PURPYR", "DNA4", "DNA", "DNA5"]
MIXED_AA = AA20 + AA20.lower() + b"XBZOU*-"
MIXED_NT = b"ACGTacgtNnUuRYKM-"


def h(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def hs(strs):
    return hashlib.sha256("\n".join(strs).encode("latin-1")).hexdigest()[:16]


def lut_of(tok):
    lut = np.full(256, -1, dtype=np.int8)
    for k, v in tok.token_decoder().items():
        for byte in v:
            lut[byte] = k
    return lut


golden = {"alphabets": {}, "kats": [], "hashes": []}

for k in KEYS:
    t = R.Tokenizer(k, bos=True, eos=True, padchar=True)
    lut = lut_of(R.Tokenizer(k))
    golden["alphabets"][k] = {
        "nchars": t.nchars(), "bos": t.bos(), "eos": t.eos(), "pad": t.pad(), "alphabet_size": t.alphabet_size(),
        "lut_hex": lut.tobytes().hex(), "lut_sha": h(lut),
        "lookup": {str(a): b for a, b in sorted(R.Tokenizer(k, bos=True, eos=True, padchar=True).lut().items())}
        if k != "BYTES" else None,
    }


def kat(key, flags, seqs, padlen, **kw):
    t = R.Tokenizer(key, **flags)
    op = kw.pop("op", "tokenize")
    rec = {"key": key, "flags": flags, "seqs": seqs, "padlen": padlen, "op": op, "kw": kw}
    if op == "tokenize":
        out = t.batch_tokenize(seqs, padlen=padlen, **kw)
        rec["out"] = out.tolist()
        rec["decoded"] = t.decode_tokens(out)
    else:
        out = t.batch_onehot_encode(seqs, padlen=padlen, **kw)
        rec["out"] = out.tolist()
    rec["dtype"] = str(out.dtype)
    golden["kats"].append(rec)


PBEOS = dict(bos=True, eos=True, padchar=True)
kat("DNA", PBEOS, ["ACGT", "GGGG"], 7, batch_first=True)            # README.md:38-44
kat("DNA", {}, ["ACGT", "GG"], 6, batch_first=True)
kat("DNA", dict(padchar=True), ["ACGT", "GG"], 6, batch_first=True)
kat("PROTEIN", PBEOS, ["", "A"], 4, batch_first=True)
kat("PROTEIN", PBEOS, ["AC"], 4, batch_first=True)
kat("DNA", {}, ["UuNnacgtACGT"], 12, batch_first=True)
kat("DNA", {}, ["GG"], 4, batch_first=True)
kat("DNA", PBEOS, ["ACGT", "GG", ""], 6, batch_first=False)
kat("SEB6", {}, ["ACDEFGHIKLMNPQRSTVWY"], 20, batch_first=True)
kat("DNA", {}, ["ACGT", "GN"], 5, op="onehot")
kat("DNA", PBEOS, ["ACGT", "GN", ""], 7, op="onehot", destchar="f")
kat("DNA5", dict(eos=True), ["ACGTNRYKM-", "acgtn"], 12, batch_first=True, destchar="i")
kat("BYTES", dict(bos=True, eos=True, padchar=True), ["Az09 ~", "\x01\x7f"], 9, batch_first=True, destchar="h")


def hcase(name, inp, key, flags, padlen, op="tokenize", mask_seed=None, decode=False, **kw):
    seed, n, lo, hi, alpha = inp
    buf, offs = gen(seed, n, lo, hi, alpha)
    seqs = as_list(buf, offs)
    t = R.Tokenizer(key, **flags)
    rec = {"name": name, "gen": [seed, n, lo, hi, alpha.decode("latin-1")], "buf_sha": h(buf), "offs_sha": h(offs),
           "key": key, "flags": flags, "padlen": padlen, "op": op, "kw": kw}
    if op == "tokenize":
        out = t.batch_tokenize(seqs, padlen=padlen, **kw)
        if decode:
            rec["decode_sha"] = hs(t.decode_tokens(out))
    else:
        if mask_seed is not None:
            m = gen_mask(mask_seed, buf.size)
            rec["mask_seed"] = mask_seed
            kw = dict(kw, mask=[m[offs[i]:offs[i + 1]] for i in range(n)])
        out = t.batch_onehot_encode(seqs, padlen=padlen, **kw)
    rec["sha"] = h(out)
    rec["shape"] = list(out.shape)
    rec["itemsize"] = out.itemsize
    golden["hashes"].append(rec)


in1 = (1, 64, 1000, 1000, b"ACGT")
in2 = (2, 256, 50, 1022, AA20)
in3 = (3, 32, 4096, 4096, b"ACGT")
in4 = (4, 512, 1, 300, MIXED_AA)
in5 = (5, 128, 0, 200, MIXED_NT)
in6 = (6, 77, 0, 130, MIXED_AA)     # odd batch size, tiny and empty sequences
hcase("G1", in1, "DNA", {}, 1024, batch_first=True)
hcase("G1s", in1, "DNA", {}, 1024, batch_first=False)
hcase("G2", in2, "PROTEIN", PBEOS, 1024, batch_first=True, decode=True)
hcase("G2s", in2, "PROTEIN", PBEOS, 1024, batch_first=False)
hcase("G2u", in2, "PROTEIN", PBEOS, 1026, batch_first=True)
hcase("G2us", in2, "PROTEIN", PBEOS, 1026, batch_first=False)
hcase("G2i", in2, "PROTEIN", PBEOS, 1024, batch_first=True, destchar="i")
hcase("G2o", in2, "PROTEIN", PBEOS, 1024, op="onehot")
hcase("G2om", in2, "PROTEIN", PBEOS, 1024, op="onehot", mask_seed=22)
hcase("G2e", in2, "PROTEIN", dict(eos=True), 1024, batch_first=True)
hcase("G3", in3, "DNA", {}, 4096, op="onehot", destchar="f")
for k in ["SEB6", "SEB8", "SEB10", "SEB14", "SEV10", "MURPHY", "LIA10", "LIB10", "DAYHOFF"]:
    hcase("in4_" + k, in4, k, PBEOS, 304, batch_first=True, decode=True)
for k in ["DNA", "DNA5", "KETO", "PURPYR", "DNAMETH"]:
    fl = dict(bos=True, eos=False, padchar=True)
    hcase("in5_" + k, in5, k, fl, 201, batch_first=True, decode=True)
    hcase("in5o_" + k, in5, k, fl, 201, op="onehot", destchar="f")
for dc in "bhilqfd":
    for bf in (True, False):
        hcase(f"in6_{dc}_{'bf' if bf else 'sf'}", in6, "PROTEIN", PBEOS, 133, batch_first=bf, destchar=dc)
    hcase(f"in6o_{dc}", in6, "SEB14", dict(bos=True, eos=True), 135, op="onehot", destchar=dc, mask_seed=66)
hcase("in6_bytes_h", in6, "BYTES", PBEOS, 140, batch_first=True, destchar="h")
hcase("in6_bytes_b", in6, "BYTES", PBEOS, 140, batch_first=False, destchar="b")

out_path = os.path.join(ROOT, "tests", "golden", "golden.json")
with open(out_path, "w") as f:
    json.dump(golden, f, indent=0, sort_keys=True)
print("wrote", out_path, os.path.getsize(out_path), "bytes;", len(golden["kats"]), "kats,", len(golden["hashes"]), "hash cases")
for r in golden["hashes"][:11]:
    print(r["name"], r["sha"], r.get("decode_sha", ""))
