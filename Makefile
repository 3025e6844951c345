# Makefile -- build the engine without Python (what a C maintainer of smvp-toolkit would run).
#   make            libsmvp_cuda.so + libsmvp_host.so + smvp-toolkit-cli under smvp-toolkit_b200/lib/
#   make oracle     the test oracle (and oracle/_ref when /root/reference is present)
#   make test       CPU test suite
# `python __graft_entry__.py` does the same through smvp-toolkit_b200/build.py (incremental, parallel).
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        ?= gcc
PKG       := smvp-toolkit_b200
LIB       := $(PKG)/lib
OBJ       := $(LIB)/obj
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DSMVP_BUILDING_LIB
CFLAGS    := -O2 -std=gnu11 -Wall -Wextra -D_XOPEN_SOURCE=700 -Iinclude -I$(PKG)/host

CU_SRCS   := $(wildcard $(PKG)/csrc/*.cu)
CU_OBJS   := $(patsubst $(PKG)/csrc/%.cu,$(OBJ)/%.o,$(CU_SRCS))
HOST_LIB  := $(filter-out %/main-cli.c,$(wildcard $(PKG)/host/*.c))

.PHONY: all oracle test clean
all: $(LIB)/libsmvp_cuda.so $(LIB)/libsmvp_host.so $(LIB)/smvp-toolkit-cli

$(OBJ)/%.o: $(PKG)/csrc/%.cu $(PKG)/csrc/common.cuh include/smvp_cuda.h include/smvp_synth.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB)/libsmvp_cuda.so: $(CU_OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $^

$(LIB)/libsmvp_host.so: $(HOST_LIB) $(PKG)/host/smvp_mmio.h $(PKG)/host/smvp_host.h
	@mkdir -p $(LIB)
	$(CC) $(CFLAGS) -fPIC -shared -o $@ $(HOST_LIB) -lm -lpthread

$(LIB)/smvp-toolkit-cli: $(PKG)/host/main-cli.c $(HOST_LIB) $(LIB)/libsmvp_cuda.so
	$(CC) $(CFLAGS) -o $@ $(PKG)/host/main-cli.c $(HOST_LIB) -L$(LIB) -lsmvp_cuda -Wl,-rpath,'$$ORIGIN' -lm -lpthread

oracle:
	$(MAKE) -C oracle all

test: all oracle
	python -m pytest tests -q -m "not gpu"

clean:
	rm -rf $(LIB)
	$(MAKE) -C oracle clean
