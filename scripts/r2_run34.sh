#!/bin/bash
timeout 120 python scripts/probe_rot_phases.py 10000000 2>&1 | tail -4
