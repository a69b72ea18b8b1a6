#!/bin/bash
# r02 evidence refresh: round script (tests, both bench arms, sweep, launch list, latency) + ncu summaries of every kernel family
bash tools/gpu_r02_round.sh
bash tools/prof_r02.sh
