timeout 800 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -3
for c in "vod clustered 16 30000" "vod uniform 16 30000" "tj4d clustered 16 30000" "tj4d uniform 16 30000" "stress clustered 16 200000"; do python scripts/quick_cfg.py $c 2>&1 | tail -1; done
for k in 1 2 4 8; do echo "== HGSF_TILE_CHUNK=$k"; for c in "vod clustered 16 2000" "tj4d clustered 16 2000" "tj4d clustered 16 30000" "vod clustered 16 10000"; do HGSF_TILE_CHUNK=$k python scripts/quick_cfg.py $c 2>&1 | tail -1; done; done
