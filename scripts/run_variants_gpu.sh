timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q 2>&1 | tail -2
for v in "" ilp2 ilp3; do
  echo "== variant ${v:-default(ilp4)}"
  if [ -n "$v" ]; then export HGSF_LIB=hgsfusion_b200/variants/$v.so; fi
  timeout 200 python scripts/quick_gpu.py 2>&1 | tail -2
  timeout 200 python scripts/quick_gpu.py 2>&1 | tail -2
done
