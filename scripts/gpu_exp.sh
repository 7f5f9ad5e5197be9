#!/bin/bash
timeout 600 python scripts/tune_r2.py 16 256,0,0 2>&1 | tail -1
