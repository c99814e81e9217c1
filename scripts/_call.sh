set -x
python -m pytest tests/test_bench_scale_parity.py -m gpu -x -q > gpurun_out/r02_scale_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_scale_parity.log
python scripts/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'project_kernel|logits_mma_kernel|point_logits_kernel|logits_prep' -c 14 -o gpurun_out/prof_r02a python scripts/prof_kernels.py > gpurun_out/prof_ncu.log 2>&1
echo done
