#!/bin/bash
timeout 120 python scripts/probe_rot100.py 2>&1 | tail -8
