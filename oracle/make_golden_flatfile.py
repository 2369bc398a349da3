"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/flatfile.json from the reference's own
FlatFile (src/fxstats.cpp, compiled into oracle/_ref by oracle/Makefile):

    make -C oracle ref && python oracle/make_golden_flatfile.py

For each FASTA/FASTQ input of oracle/fastx_cases.py (given both plain and, for every fifth
case, gzip-compressed) the fixture records the input text, the exact bytes of the flat file
the reference wrote, and what its reader reports (nseqs, seq_offset, maxseqlen, getstats)."""
import gzip
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import load_ref  # noqa: E402
from oracle.fastx_cases import cases  # noqa: E402

R = load_ref()
assert R is not None and hasattr(R, "FlatFile"), "build the reference first: make -C oracle ref"

out = []
with tempfile.TemporaryDirectory() as d:
    for i, data in enumerate(cases()):
        gz = i % 5 == 4
        src = os.path.join(d, f"c{i}.fa" + (".gz" if gz else ""))
        with (gzip.open if gz else open)(src, "wb") as f:
            f.write(data)
        dst = os.path.join(d, f"c{i}.ff")
        made = R.FlatFile(src, dst)      # writes the file (accessing this object is what crashes upstream)
        del made
        ff = R.FlatFile(dst)
        out.append({"input": data.decode("latin-1"), "gz": gz, "ff_hex": open(dst, "rb").read().hex(),
                    "nseqs": ff.nseqs(), "seq_offset": ff.seq_offset(), "maxseqlen": ff.maxseqlen,
                    "lens": R.getstats([src])[0].tolist(),
                    "seqs": [bytes(x).decode("latin-1") for x in ff.access(0, ff.nseqs())]})
path = os.path.join(ROOT, "tests", "golden", "flatfile.json")
with open(path, "w") as f:
    json.dump(out, f, indent=0)
print("wrote", path, os.path.getsize(path), "bytes,", len(out), "cases,", sum(c["nseqs"] for c in out), "sequences")
