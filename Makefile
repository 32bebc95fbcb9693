# xmp_b200 build: host C library, sm_100a CUDA library (C ABI), the fastdnamp_b200 CLI, and the
# test-only oracle.  Everything is built in-tree so that the .so files travel with the snapshot.
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v --use_fast_math
CFLAGS    := -O3 -mpopcnt -std=c99 -D_POSIX_C_SOURCE=200809L -Wall -Wextra -fPIC

CUDA_SRCS := xmp_b200/csrc/api.cu xmp_b200/csrc/kernels.cu xmp_b200/csrc/search.cu xmp_b200/csrc/exchange.cu xmp_b200/csrc/microbench.cu
CUDA_HDRS := xmp_b200/csrc/ctx.h xmp_b200/csrc/fitch_dev.cuh xmp_b200/csrc/child_walk.cuh xmp_b200/csrc/walk_prog.h xmp_b200/csrc/tree_book.h xmp_b200/csrc/exchange_proto.h include/xmp_cuda.h
HOST_SRCS := host/phylip.c host/prep.c host/trees.c host/partbound.c

.PHONY: all host cuda cli oracle ref checks clean
all: host cuda cli oracle checks

host: host/libxmp_host.so
host/libxmp_host.so: $(HOST_SRCS) host/xmp_host.h
	$(CC) $(CFLAGS) -shared -o $@ $(HOST_SRCS) -lpthread

cuda: xmp_b200/libxmp_b200.so
xmp_b200/libxmp_b200.so: $(CUDA_SRCS) $(CUDA_HDRS)
	$(NVCC) $(NVFLAGS) -shared -cudart static -o $@ $(CUDA_SRCS) 2> xmp_b200/csrc/ptxas.log || (cat xmp_b200/csrc/ptxas.log; exit 1)

cli: fastdnamp_b200
fastdnamp_b200: host/fastdnamp.c host/libxmp_host.so xmp_b200/libxmp_b200.so
	$(CC) $(CFLAGS) -fPIE -o $@ host/fastdnamp.c -Ihost -Iinclude -Lhost -Lxmp_b200 -lxmp_host -lxmp_b200 \
	  -Wl,-rpath,'$$ORIGIN/host' -Wl,-rpath,'$$ORIGIN/xmp_b200' -lpthread

oracle:
	$(MAKE) -C oracle oracle

# stand-alone GPU check programs (test infrastructure: they link the oracle as the checker)
checks: tests/sweep_m_check
tests/sweep_m_check: tests/sweep_m_check.cu xmp_b200/libxmp_b200.so include/xmp_cuda.h | oracle
	$(NVCC) -O2 -std=c++17 $(ARCH) -Iinclude -o $@ tests/sweep_m_check.cu -Lxmp_b200 -lxmp_b200 -Loracle -loracle \
	  -Xlinker -rpath,'$$ORIGIN/../xmp_b200' -Xlinker -rpath,'$$ORIGIN/../oracle'
ref:
	$(MAKE) -C oracle ref

clean:
	rm -f host/libxmp_host.so xmp_b200/libxmp_b200.so fastdnamp_b200 xmp_b200/csrc/ptxas.log tests/sweep_m_check
	$(MAKE) -C oracle clean
