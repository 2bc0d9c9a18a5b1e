# compute-sanitizer over the kernels added / rewritten in round 2: TMA-staged BC4/BC5 kernel (mbarriers), rolled setup / RangeFit front end,
# mip_chain_kernel + texture-group encode, decode pipelines (memcheck, racecheck, synccheck) on small GPU tests.
# --num-cuda-barriers: the TMA kernel keeps 48 mbarriers per CTA (16 warps x 3 stages); synccheck's default tracking table overflows.
echo "compute-sanitizer $(compute-sanitizer --version | tail -1), $(nvidia-smi --query-gpu=name --format=csv,noheader), round 2"
for tool in memcheck racecheck synccheck; do
  echo "$tool: test_gpu_alpha_lattice.py -k 'tma_staged and not 4096' + test_gpu_parity.py -k 'batch_groups or mip_chain_kernel or mipchain or rangefit_blocks or decompress_multi or kat'"
  if [ $tool = synccheck ]; then
    # (with the enlarged barrier table the tool fails to track the launches of other kernels: two runs, the second one with the TMA kernel off)
    compute-sanitizer --tool synccheck --num-cuda-barriers 65536 python -m pytest tests/test_gpu_alpha_lattice.py -q -x -k "tma_staged and not 4096" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|rror" | cut -c1-200 | head -6
    TXP_ALPHA_TMA=0 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -q -x -k "batch_groups or mip_chain_kernel or mipchain or rangefit_blocks or decompress_multi or kat" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|rror" | cut -c1-200 | head -6
  else
    compute-sanitizer --tool $tool python -m pytest tests/test_gpu_alpha_lattice.py tests/test_gpu_parity.py -q -x -k "(tma_staged and not 4096) or batch_groups or mip_chain_kernel or mipchain or rangefit_blocks or decompress_multi or kat" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|rror" | cut -c1-200 | head -6
  fi
done
