compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -v -x -k "batch_groups or mip_chain_kernel or mipchain or rangefit_blocks or decompress_multi" > gpurun_out/race_set.txt 2>&1
grep -c "Race reported" gpurun_out/race_set.txt
grep -n "Race reported\|PASSED\|FAILED" gpurun_out/race_set.txt | grep -B1 "Race reported" | head -20
grep -A14 "Race reported" gpurun_out/race_set.txt | head -45
compute-sanitizer --tool synccheck --num-cuda-barriers 65536 python -m pytest tests/test_gpu_alpha_lattice.py -q -x -k "tma_staged and not 4096" > gpurun_out/sync_tma.txt 2>&1
grep "=========\|passed\|failed" gpurun_out/sync_tma.txt | head -20
