cp biscuit_b200/libbiscuit_b200.so /tmp/lib_default.so
for v in stall stall_noepi stall_noprod; do
  cp profiles/variants/lib_$v.so biscuit_b200/libbiscuit_b200.so
  echo "== $v"
  timeout 200 python profiles/run_predict.py 1024 512 2>&1 | grep "sepmid stall" | tail -12
done
cp /tmp/lib_default.so biscuit_b200/libbiscuit_b200.so
