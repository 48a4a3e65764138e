#!/bin/bash
python tools/dbg_bulk.py 2>&1 | tail -30
