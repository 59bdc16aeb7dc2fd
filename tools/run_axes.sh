python tools/run_case.py > gpurun_out/ax_auto.log 2>&1; echo rc=$? >> gpurun_out/ax_auto.log
CX_VV_CHUNK=777 python tools/run_case.py > gpurun_out/ax_chunk.log 2>&1; echo rc=$? >> gpurun_out/ax_chunk.log
CX_VV_SCAN=1 python tools/run_case.py > gpurun_out/ax_scan.log 2>&1; echo rc=$? >> gpurun_out/ax_scan.log
CX_VV_PLANE_CAP=100 python tools/run_case.py > gpurun_out/ax_overflow.log 2>&1; echo rc=$? >> gpurun_out/ax_overflow.log
for ax in 0 1 2; do CX_VV_AXIS=$ax python tools/run_case.py > gpurun_out/ax_$ax.log 2>&1; echo rc=$? >> gpurun_out/ax_$ax.log; done
python tools/time_vv.py 2
python tools/time_vv.py 4 300
python tools/time_vv.py 3 500
python tools/exp_vv_axis.py
