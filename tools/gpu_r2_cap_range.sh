cap() { name=$1; k=$2; s=$3; units=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o /tmp/$name -f "$@" > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep $units gpurun_out/${name}_lines.txt > gpurun_out/${name}_summary.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null; rm -f gpurun_out/${name}_lines.txt; }
cap ncu_range_r02 range_encode 1 1048576 python tools/prof_range.py
head -22 gpurun_out/ncu_range_r02_summary.txt; tail -14 gpurun_out/ncu_range_r02_summary.txt
