#!/bin/bash
# final visit of a round: full gpu suite, smoke, micro-benchmarks, headline bench, then the ncu evidence
R=${1:-r1z}
bash tools/gpu_round.sh $R noncu
bash tools/gpu_prof.sh $R
