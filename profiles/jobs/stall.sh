cp biscuit_b200/libbiscuit_b200.so /tmp/lib_default.so
for f in profiles/variants/lib_f_*.so; do
  v=$(basename $f .so)
  cp $f biscuit_b200/libbiscuit_b200.so
  echo "== $v"
  timeout 200 python profiles/run_predict.py 1024 512 2>&1 | grep "sepmid stall cluster 0" | tail -6
done
cp /tmp/lib_default.so biscuit_b200/libbiscuit_b200.so
