set -x
python tools/run_case.py > gpurun_out/ax_auto.log 2>&1; echo rc=$? >> gpurun_out/ax_auto.log
for ax in 0 1 2; do CX_VV_AXIS=$ax python tools/run_case.py > gpurun_out/ax_$ax.log 2>&1; echo rc=$? >> gpurun_out/ax_$ax.log; done
python tools/exp_vv_axis.py > gpurun_out/exp_axis2.log 2>&1
python bench.py --config 2 --steps 5 --warmup 3 > gpurun_out/b2_ax.log 2> gpurun_out/b2_ax.err
tail -2 gpurun_out/ax_*.log; cat gpurun_out/exp_axis2.log
