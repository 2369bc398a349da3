#!/bin/bash
# ncu --set full of the SURVEY 8(f) kernels (K5 one-hot (B,C,L), K6 embedding gather, K7 augmentation) on the bench's f_rows cases
O=gpurun_out/${1:-frows}
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'onehot_bcl|embed_kernel|augment_kernel' -s 6 -c 6 -o $O/prof_frows \
    python bench.py --steps 3 --warmup 3 --smi off --sections value,frows > $O/prof_frows.log 2>&1; echo "ncu rc=$?"
