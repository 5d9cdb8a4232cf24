timeout 600 python -m pytest tests/test_i2i.py -m gpu -q --timeout 120 2>&1 | tail -3
for i in 1 2; do timeout 300 python tests/gpu_bringup.py --stage i2i_time 2>&1 | grep "fp16 ours\|speed-up"; done
INNFER_I2I_DEBUG=7 timeout 300 python tests/gpu_bringup.py --stage i2i_time 2>&1 | grep "fp16 ours"
