#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_variants.py -m gpu -q -k polarised 2>&1 | tail -12
