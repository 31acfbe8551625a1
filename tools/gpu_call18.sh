#!/bin/bash
python tools/launch_overhead.py C2
python tools/launch_overhead.py C3
