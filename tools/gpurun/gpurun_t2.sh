timeout 900 python -m pytest tests -m gpu -x -q -k "shards_over_devices or shard_invariance or devices" 2>&1 | tail -4
