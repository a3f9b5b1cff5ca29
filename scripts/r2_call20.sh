#!/bin/bash
timeout 600 python scripts/mega_cache_check.py 60 2>&1 | tail -30
