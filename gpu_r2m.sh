#!/bin/bash
./gpu_prof_r2.sh
du -sh gpurun_out
