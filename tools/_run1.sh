tools/gpu_round.sh r2k tests smoke bench ref ncu
