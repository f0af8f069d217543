#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_train_loop.py -m gpu -q 2>&1 | tail -15
