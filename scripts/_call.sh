set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_canaries.py -m gpu -q -x -k "logits or canar or ensemble or fused" > gpurun_out/gpu_tests51.log 2>&1
XM3D_SO=xmask3d_b200/libxm3d_prev.so timeout 300 python scripts/time_point_logits.py > gpurun_out/time_pl51p.log 2>&1
timeout 300 python scripts/time_point_logits.py > gpurun_out/time_pl51n.log 2>&1
echo done
