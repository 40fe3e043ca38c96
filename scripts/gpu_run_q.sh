python -m pytest tests/test_quirks.py -m gpu -q -k "quirk04 or quirk05 or quirk13" 2>&1 | grep -E "^E|passed|failed|assert" | head -40
