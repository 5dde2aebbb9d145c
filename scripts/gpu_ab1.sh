#!/bin/bash
# A/B of library variants on one GPU, no test suite
P=epidemicsimulator_b200
V=""
for name in "$@"; do V="$V ESIM_B200_LIB=$P/libesim_b200$name.so"; done
python scripts/kstep_ab.py --steps 240 $V
python scripts/kstep_ab.py --steps 120 --areas 27500 --cross 0.9 $V
