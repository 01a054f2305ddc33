#!/bin/bash
./gpu_multi2.sh 8
