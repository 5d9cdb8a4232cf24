for v in 0 1; do echo "INNFER_L2_PERSIST=$v"; INNFER_L2_PERSIST=$v INNFER_MB=95,48,32,24,16,12 timeout 300 python tests/gpu_bringup.py --stage time 2>&1 | grep "iter=[12]\|carve"; done
