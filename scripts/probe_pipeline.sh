timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== default lib"; timeout 200 python scripts/pipeline_probe.py 2>&1 | tail -4
export HGSF_LIB=hgsfusion_b200/variants/exp.so
echo "== exp, emit 2/SM"; HGSF_EXTRA_SMEM=26000 timeout 200 python scripts/pipeline_probe.py 2>&1 | tail -4
echo "== exp, emit 2/SM, front 2/SM"; HGSF_EXTRA_SMEM=26000 HGSF_FRONT_CTAS=2 timeout 200 python scripts/pipeline_probe.py 2>&1 | tail -4
echo "== exp, emit 2/SM, front 1/SM"; HGSF_EXTRA_SMEM=26000 HGSF_FRONT_CTAS=1 timeout 200 python scripts/pipeline_probe.py 2>&1 | tail -4
echo "== default emit 3/SM, front 1/SM"; HGSF_FRONT_CTAS=1 timeout 200 python scripts/pipeline_probe.py 2>&1 | tail -4
