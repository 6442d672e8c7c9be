#!/bin/bash
timeout 600 python tools/dbg_dimitrov.py PbFlorentSeq2 2>&1 | tail -60
