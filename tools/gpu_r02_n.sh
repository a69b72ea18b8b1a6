#!/bin/bash
# r02 refresh after the canon change: the round script + ncu of the Solinas kernels only
bash tools/gpu_r02_round.sh
ONLY='^ntt64s' bash tools/prof_r02.sh
