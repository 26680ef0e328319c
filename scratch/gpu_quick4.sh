#!/bin/bash
timeout 120 python scratch/prof_vec.py 100 2>&1 | tail -1
timeout 120 python scratch/prof_vec.py 100 row 2>&1 | tail -1
