timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r01l_launches_train_step.csv python tools/train_once.py > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r01l_launches_train_step.csv 12
