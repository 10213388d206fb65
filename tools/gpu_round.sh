python -m pytest tests/test_gpu_parity.py -q -x -k "dense or frame_batch" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1v_cfg5_launches.csv python tools/cfg5_cprofile.py > gpurun_out/ncu_cfg5.log 2>&1
grep -E "k_chol_small" gpurun_out/r1v_cfg5_launches.csv | tail -2 | cut -d, -f5,15-
python bench.py --workload cfg5 2>/dev/null | tail -1 | cut -c1-200
