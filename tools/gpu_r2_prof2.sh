# round 2: ncu --set full summaries of the remaining kernel families (iterative lane search, warp-per-block search, decode, mip chain)
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, units, target...
  name=$1; k=$2; s=$3; units=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o /tmp/$name -f "$@" > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep $units gpurun_out/${name}_lines.txt > gpurun_out/${name}_summary.txt 2>&1
  head -20 gpurun_out/${name}_summary.txt | cut -c1-140
}
cap ncu_lane_iter3_r02 cluster_lane_iter 2 1048576 python tools/prof_iter.py
cap ncu_lane_iter4_r02 cluster_lane_iter 3 1048576 python tools/prof_iter.py
cap ncu_decode_bc3_r02 decode_kernel 3 4194304 python tools/prof_misc.py
cap ncu_warp_search_r02 colour_search_kernel 1 65536 python tools/prof_small.py
cap ncu_mip_chain_r02 mip_chain_kernel 1 8 python tools/prof_small.py
rm -f gpurun_out/*_lines.txt
