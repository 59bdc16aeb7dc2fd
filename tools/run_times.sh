python tools/time_vv.py 2
python tools/time_vv.py 4 300
python tools/exp_vv_axis.py
