# A/B of the BC4/BC5 encoder variants on one B200 (TXP_ALPHA_VARIANT / TXP_ALPHA_STAGED are tuning knobs of txp_api.cu)
python -m pytest tests/test_gpu_alpha_lattice.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for cfg in ${CFGS:-"0 1" "0 0" "6 1" "8 1"}; do set -- $cfg; echo "variant $1 staged $2"; TXP_ALPHA_VARIANT=$1 TXP_ALPHA_STAGED=$2 python tools/bench_extra.py --cases bc4,bc5 --reps 5 2>&1 | tail -2 | cut -c1-120; done
