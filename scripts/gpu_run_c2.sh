python -m pytest tests/test_roq_basis.py -x -q -s 2>&1 | tail -25
