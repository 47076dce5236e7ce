#!/bin/bash
nproc; lscpu | grep -E "Socket|NUMA|Model name" ; nvidia-smi topo -m 2>/dev/null | head -14; free -g | head -2
