# launch list of one warm micro-batch (times + DRAM bytes per launch), then full captures of the kernels changed this round
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2v.csv python profiles/run_predict.py 1024 512 > gpurun_out/launch_run.log 2>&1
tail -2 gpurun_out/launch_run.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1_tc -s 1 -c 1 -o gpurun_out/conv1tc_r2v -f python profiles/run_predict.py 1024 512 > gpurun_out/conv1_ncu.log 2>&1
tail -1 gpurun_out/conv1_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_is -s 1 -c 1 -o gpurun_out/conv2_r2v -f python profiles/run_predict.py 1024 512 > gpurun_out/conv2_ncu.log 2>&1
tail -1 gpurun_out/conv2_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sepconv_mid -s 25 -c 3 -o gpurun_out/sepmid_r2v -f python profiles/run_predict.py 1024 512 > gpurun_out/sepmid_ncu.log 2>&1
tail -1 gpurun_out/sepmid_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sepconv2d_fused -s 4 -c 4 -o gpurun_out/sep2d_r2v -f python profiles/run_predict.py 1024 512 > gpurun_out/sep2d_ncu.log 2>&1
tail -1 gpurun_out/sep2d_ncu.log
