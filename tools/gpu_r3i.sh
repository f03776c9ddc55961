# last validation of the round: the GPU tests behind the one that hung in the previous call (its cause: a launch site of
# segment_offsets_kernel that kept the one-thread-per-cluster grid), in two short calls
mkdir -p gpurun_out
timeout ${1:-100} python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "${2:-kmeanspp or graph or cta_pair or adaptive or staged}" > gpurun_out/r3j_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r3j_pytest.txt
