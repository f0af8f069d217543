#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_pretrain.py -m gpu -q 2>&1 | tail -25
