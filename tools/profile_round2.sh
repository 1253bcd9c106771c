#!/bin/bash
# Round-2 ncu captures (run on the GPU box, one GPU): `--set full` of every kernel the bench times, summarised to
# gpurun_out/r02_ncu_<name>.json by tools/ncu_summary.py (the .ncu-rep files stay on the box: they exceed the copy-back
# limit), plus the launch list of one bench-shaped pass.  Usage: bash tools/profile_round2.sh [names...]
set -u
OUT=gpurun_out
mkdir -p $OUT
cap() {   # name, kernel regex, launches to skip, command...
    local name=$1 regex=$2 skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip $skip --launch-count 1 -o /tmp/r02_$name -f "$@" > $OUT/r02_ncu_$name.log 2>&1
    ncu -i /tmp/r02_$name.ncu-rep --page raw --csv > /tmp/r02_${name}_raw.csv 2>/dev/null
    ncu -i /tmp/r02_$name.ncu-rep --page source --csv > /tmp/r02_${name}_source.csv 2>/dev/null
    python tools/ncu_summary.py /tmp/r02_${name}_raw.csv /tmp/r02_${name}_source.csv > $OUT/r02_ncu_$name.json 2>>$OUT/r02_ncu_$name.log
    rm -f /tmp/r02_$name.ncu-rep
}
want() { [ $# -eq 0 ] && return 0; for w in "$@"; do [ "$w" = "$NAME" ] && return 0; done; return 1; }
for NAME in launches features_2048 features_16384 onegnn_tc front_end_16384 col_argmin_16384 min_trick_16384 front_end_2048 k_solve; do
    want "$@" || continue
    case $NAME in
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r02_launches_profile_target.csv python tools/profile_target.py --reps 1 --batch 64 > $OUT/r02_launches.log 2>&1 ;;
    features_2048) cap $NAME k_row_features_group 1 python tools/profile_features_group.py 2048 64 ;;
    features_16384) cap $NAME k_row_features_group 1 python tools/profile_features_group.py 16384 1 ;;
    onegnn_tc) cap $NAME k_onegnn_tc 1 python tools/profile_target.py --what pipeline --batch 64 --reps 1 ;;
    front_end_16384) cap $NAME k_front_end 1 python tools/profile_target.py --what dense --reps 1 ;;
    col_argmin_16384) cap $NAME k_col_argmin_partial 1 python tools/profile_target.py --what dense --reps 1 ;;
    min_trick_16384) cap $NAME k_min_trick_partial 1 python tools/profile_target.py --what dense --reps 1 ;;
    front_end_2048) cap $NAME k_front_end 1 python tools/profile_target.py --what pipeline --batch 64 --reps 1 ;;
    k_solve) cap $NAME k_solve 1 python tools/profile_target.py --what pipeline --batch 64 --reps 1 ;;
    esac
done
ls -la $OUT | tail -12
