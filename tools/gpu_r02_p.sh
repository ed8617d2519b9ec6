#!/bin/bash
exec bash tools/gpu_r02_records.sh "$@"
