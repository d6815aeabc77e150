python -m pytest tests -m gpu -x -q -k channel_concat 2>&1 | tail -40
