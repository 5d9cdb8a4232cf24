timeout 600 python -m pytest tests/test_i2i.py -m gpu -q --timeout 120 2>&1 | tail -3
for g in 1 0; do echo "INNFER_I2I_GRAPH=$g"; INNFER_I2I_GRAPH=$g timeout 300 python tests/gpu_bringup.py --stage i2i_time 2>&1 | grep "fp16 ours\|speed-up"; done
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 10000000 --show-backtrace no python -m pytest tests/test_gpu_parity.py -q -m gpu -k "conv_block_wide_layout and 192" 2>&1 | python tools/racecheck_fold.py | cut -c1-300 | tail -30
