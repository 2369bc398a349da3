"""TEST INFRASTRUCTURE ONLY.  FASTA/FASTQ inputs for the FlatFile parity tests: hand-written
edge cases plus a seeded random generator (well-formed records, FASTQ with broken quality
strings, CRLF, blank lines, truncation, raw token soup)."""
import random

HAND = [
    b"",
    b">",
    b">\n",
    b">a",
    b">a\n",
    b">a\nACGT",
    b">a\nACGT\n",
    b">s1 comment\nACGT\nAC\n\n>s2\n\nGG\r\nTT\n>s3\n>s4\nA",
    b"junk before\n>a desc\nAC GT\n\tNN\n",
    b">a\r\nACGT\r\nAC\r\n>b\r\n\r\nGG\r\n",
    b">a\n\r\n>b\n\r",
    b">a\nAC>GT\n>b\nA@C\n",
    b"@r1\nACGT\n+\nIIII\n@r2\nGG\n+r2\n!!\n",
    b"@r1\nACGT\n+\nII\nII\n@r2\nGG\n+\n@>\n",
    b"@r1\nAC\nGT\n+\nIIII\n",
    b"@r1\nACGT\n+\nIII\n@r2\nGG\n+\n!!\n",
    b"@r1\nACGT\n+\nIIIII\n@r2\nGG\n+\n!!\n",
    b"@r1\nACGT\n+",
    b"@r1\nACGT\n+\n",
    b"@r1\n\n+\n\n@r2\nA\n+\nI\n",
    b"@r1\r\nACGT\r\n+\r\nIIII\r\n@r2\r\nGG\r\n+\r\n!!\r\n",
    b">a\nACGT\n@b\nGG\n+\nII\n>c\nTT\n",
    b">prot\nMKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQAPILSRVGDGTQDNLSGAEKAVQVKVKALPDAQFEVV\nHSLAKWKR\n>p2\nMK*\n",
]


def random_fastx(rng):
    pieces = [b">", b"@", b"+", b"\n", b"\r\n", b"\r", b" ", b"\t", b"ACGT", b"A", b"NNNN", b"acgu", b"IIII", b"!",
              b">x y z", b"@r1", b"+\n", b"\n\n", b"*", b"-"]
    mode = rng.random()
    out = []
    if mode < 0.4:
        for _ in range(rng.randint(0, 8)):
            out.append(b">" + rng.choice([b"id", b"id desc", b"", b"id\tdesc"]) + rng.choice([b"\n", b"\r\n"]))
            for _ in range(rng.randint(0, 4)):
                out.append(b"".join(rng.choice([b"A", b"C", b"G", b"T", b"N", b" "]) for _ in range(rng.randint(0, 30)))
                           + rng.choice([b"\n", b"\r\n", b"\n\n"]))
        if out and rng.random() < 0.5:
            out[-1] = out[-1].rstrip(b"\r\n")
    elif mode < 0.75:
        for _ in range(rng.randint(0, 8)):
            n = rng.randint(0, 40)
            s = bytes(rng.choice(b"ACGTN") for _ in range(n))
            q = bytes(rng.choice(b"!#IJ@>+5") for _ in range(n if rng.random() < 0.85 else rng.randint(0, 45)))
            nl = rng.choice([b"\n", b"\r\n"])
            if rng.random() < 0.2 and n > 4:
                k = rng.randint(1, n - 1)
                s = s[:k] + nl + s[k:]
            if rng.random() < 0.2 and len(q) > 4:
                k = rng.randint(1, len(q) - 1)
                q = q[:k] + nl + q[k:]
            out.append(b"@r" + nl + s + nl + b"+" + rng.choice([b"", b"r"]) + nl + q + nl)
        if out and rng.random() < 0.3:
            out[-1] = out[-1][:rng.randint(0, len(out[-1]))]
    else:
        out = [rng.choice(pieces) for _ in range(rng.randint(0, 40))]
    return b"".join(out)


def cases(seed=20261017, nrandom=120):
    rng = random.Random(seed)
    return list(HAND) + [random_fastx(rng) for _ in range(nrandom)]
