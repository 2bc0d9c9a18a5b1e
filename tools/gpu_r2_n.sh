# A/B of the regular-path variants of the BC4/BC5 lattice kernel (tools/micro/ab_bin/alpha_ab_*: same harness, different -D flags)
for n in base err pack mm2 ep all; do ./tools/micro/ab_bin/alpha_ab_$n $n 0 2>&1 | grep TMA | head -4; done | tee gpurun_out/alpha_ab_r02n.txt
for n in base all; do ./tools/micro/ab_bin/alpha_ab_$n ${n}_smooth 1 2>&1 | grep TMA | head -4; done | tee -a gpurun_out/alpha_ab_r02n.txt
