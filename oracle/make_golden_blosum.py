"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/blosum.json from the reference's own
bioseq/blosum.py (imported from /root/reference in the build container; it is pure numpy):

  normrows   the 21 x 20 substitution probabilities the reference samples from (blosum.py:41-43)
  order      the amino-acid column order (blosum.py:36, :44)
  x_row      index of the row used for unknown residues (blosum.py:49, :60)
  sample     20000 draws of the reference's own `substitute` for three residues (its numpy
             generator), as counts -- a check that the table is what the sampler really uses

Run:  python oracle/make_golden_blosum.py   (needs /root/reference)
"""
import importlib.util
import json
import os
from collections import Counter

REF = os.environ.get("BSQ_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "blosum.json")


def main():
    spec = importlib.util.spec_from_file_location("ref_blosum", os.path.join(REF, "bioseq", "blosum.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    order = "".join(m.aa_array.tolist())
    doc = {
        "source": "bioseq/blosum.py (normrows, aa_array, probdict, substitute)",
        "order": order,
        "true_aas": m.true_aas,
        "x_row": m.true_aas.index("X"),
        "normrows": [[float(x) for x in row] for row in m.normrows],
        "unknown_uses_x_row": bool((m.probdict.get("b", m.default_transitions) == m.normrows[m.true_aas.index("X")]).all()),
        "sample": {aa: dict(Counter(m.substitute(aa, size=20000).tolist())) for aa in "HKW"},
    }
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=0)
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
