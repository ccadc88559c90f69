timeout 300 python -m pytest tests -m gpu -k "not tc" -q -p no:cacheprovider 2>&1 | tail -6
timeout 400 python -m pytest tests -m gpu -k "tc" -q -p no:cacheprovider 2>&1 | grep -v "^E  " | tail -8
