# ncu --set full of the fused middle-flow kernel (three consecutive launches: sepconv2, sepconv3 (+residual), sepconv1)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sepconv_mid -s 1 -c 3 -o gpurun_out/sepmid_r2c -f python profiles/run_predict.py 512 512 > gpurun_out/sepmid_ncu.log 2>&1
tail -3 gpurun_out/sepmid_ncu.log
