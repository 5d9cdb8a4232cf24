#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02e_pytest.log 2>&1; tail -3 gpurun_out/r02e_pytest.log
for v in 0 1; do
INNFER_ROWS_IDT=$v ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv_rows -s 0 -c 45 --csv --log-file gpurun_out/r02e_c5_$v.csv python tests/gpu_bringup.py --stage prof > /dev/null 2>&1
echo "== INNFER_ROWS_IDT=$v"; python tools/ncu_seq.py gpurun_out/r02e_c5_$v.csv 0 0
done
python tools/ncu_seq.py gpurun_out/r02e_c5_1.csv 0 16 | tail -16
for v in 0 1 0 1; do echo "== INNFER_ROWS_IDT=$v"; INNFER_ROWS_IDT=$v INNFER_MB=95 python tests/gpu_bringup.py --stage time 2>&1 | grep "time 1080p"; done
