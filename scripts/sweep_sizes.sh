for A in 2800 11300 45200; do
python scripts/profile_steps.py --areas $A --steps 24 > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:k_ -s 72 -c 48 --csv --log-file gpurun_out/sw_$A.csv python scripts/profile_steps.py --areas $A --steps 24 > gpurun_out/sw_$A.log 2>&1
python scripts/profile_steps.py --areas $A --steps 1200 --skip 48
done
