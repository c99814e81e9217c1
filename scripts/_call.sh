set -x
XM3D_SO=xmask3d_b200/libxm3d_prev.so timeout 300 python scripts/time_point_logits.py > gpurun_out/time_pl47p.log 2>&1
timeout 300 python scripts/time_point_logits.py > gpurun_out/time_pl47n.log 2>&1
XM3D_SO=xmask3d_b200/libxm3d_prev.so timeout 300 python scripts/time_logits.py >> gpurun_out/time_pl47p.log 2>&1
timeout 300 python scripts/time_logits.py >> gpurun_out/time_pl47n.log 2>&1
echo done
