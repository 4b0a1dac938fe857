#!/bin/bash
cd "$(dirname "$0")/.."
python tools/latency.py vitl14 vits14 2>&1 | grep -v dino_model | tail -6
DINO_B200_GRAPH=0 python tools/latency.py vitl14 2>&1 | grep "batch 1" 
