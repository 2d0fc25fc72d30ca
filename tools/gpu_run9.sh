#!/bin/bash
# one B200: where the time of one rank's part build goes (C4, part 0 of 8 / of 2)
set -u
O=gpurun_out
python tools/part_build_time.py --parts 8 > $O/part_build.log 2>&1
python tools/part_build_time.py --parts 2 >> $O/part_build.log 2>&1
B200_NO_GRAPH=1 python tools/part_build_time.py --parts 8 >> $O/part_build.log 2>&1
B200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/part_build_launches.csv \
   python tools/part_build_time.py --parts 8 --reps 1 > $O/ncu_part_build.log 2>&1
cat $O/part_build.log
