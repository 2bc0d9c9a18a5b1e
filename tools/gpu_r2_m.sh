# round 2, evidence call: smoke, full GPU suite, headline bench with configs, launch list, ncu --set full of the four kernel families
# (.ncu-rep files are summarised on the box with tools/ncu_summary.py and deleted: gpurun_out/ is capped at 64 MiB)
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; tail -c 400 gpurun_out/bench_r02_n1.err; cut -c1-400 gpurun_out/bench_r02_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/bench_under_ncu.log 2>&1
cap() {  # name, kernel regex, skip, units, target...
  name=$1; k=$2; s=$3; units=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o /tmp/$name -f "$@" > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep $units gpurun_out/${name}_lines.txt > gpurun_out/${name}_summary.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
}
cap ncu_lane_r02 cluster_lane_kernel 1 4194304 python tools/prof_lane.py 8192
cap ncu_setup_r02 cluster_setup_sorted 1 4194304 python tools/prof_lane.py 8192
cap ncu_setup_smooth_r02 cluster_setup_sorted 1 4194304 python tools/prof_lane.py 8192 smooth
cap ncu_lane_smooth_r02 cluster_lane_kernel 1 4194304 python tools/prof_lane.py 8192 smooth
cap ncu_range_r02 range_encode 1 1048576 python tools/prof_range.py
cap ncu_alpha_bc4_r02 alpha_lattice 1 131072 python tools/prof_alpha.py
cap ncu_alpha_bc5_r02 alpha_lattice 4 131072 python tools/prof_alpha.py
du -sh gpurun_out; ls -la gpurun_out
