set -x
for c in 25 50; do
XM3D_CARVE=$c XM3D_SO=xmask3d_b200/libxm3d_dbg.so timeout 600 python scripts/exp_proj_overlap.py > gpurun_out/exp_proj_overlap_$c.log 2>&1; echo "rc=$?" >> gpurun_out/exp_proj_overlap_$c.log
done
echo done
