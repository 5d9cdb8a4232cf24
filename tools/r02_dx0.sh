for v in 0 1 0 1; do echo "INNFER_ROWS_DX0=$v"; INNFER_ROWS_DX0=$v INNFER_MB=95 timeout 300 python tests/gpu_bringup.py --stage time 2>&1 | grep "iter=[12]"; done
