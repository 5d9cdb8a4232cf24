timeout 600 python -m pytest tests/test_i2i.py -m gpu -q --timeout 120 2>&1 | tail -3
for g in 1 0 1 0; do echo "INNFER_I2I_HALO_SPLIT=$g"; INNFER_I2I_HALO_SPLIT=$g timeout 300 python tests/gpu_bringup.py --stage i2i_time 2>&1 | grep "fp16 ours"; done
