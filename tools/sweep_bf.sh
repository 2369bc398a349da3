#!/bin/bash
python tools/sweep_bf.py 2>&1 | tail -1
PADLEN=1026 python tools/sweep_bf.py 2>&1 | tail -1
PADLEN=656 python tools/sweep_bf.py 2>&1 | tail -1
PADLEN=652 python tools/sweep_bf.py 2>&1 | tail -1
