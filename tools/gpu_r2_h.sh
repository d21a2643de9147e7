#!/bin/bash
export PYTHONPATH=$PWD
timeout 300 python tools/minwidth_probe.py
