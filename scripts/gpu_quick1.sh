#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_steps.py --steps 1920 --skip 48 --repeat 2 2>&1 | tail -2
ESIM_KTRACE=1 python scripts/profile_steps.py --steps 960 --skip 24 2>&1 | tail -8
python scripts/profile_steps.py --steps 960 --skip 48 --areas 27500 --cross 0.9 --repeat 2 2>&1 | tail -2
