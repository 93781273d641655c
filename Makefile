# Builds the sm_100a engine (nu_nerf_b200/libnunerf_b200.so) and the CPU oracle (oracle/_build/liboracle.so).
NVCC ?= nvcc
CC ?= gcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function -Xcompiler -Wno-unknown-pragmas $(EXTRA)
CSRC := nu_nerf_b200/csrc
OBJDIR := build/obj
SRCS := gemm.cu linear.cu chain.cu chain_ts.cu probe.cu weights.cu sampling.cu composite.cu field.cu bvh.cu mcubes.cu shell.cu
OBJS := $(SRCS:%.cu=$(OBJDIR)/%.o)
LIB := nu_nerf_b200/libnunerf_b200.so
ORACLE := oracle/_build/liboracle.so

all: $(LIB) $(ORACLE)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/ptx.cuh $(CSRC)/tma_host.cuh $(CSRC)/pointwise.cuh $(CSRC)/chain_common.cuh include/nunerf.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

$(ORACLE): oracle/sampling_oracle.c
	@mkdir -p oracle/_build
	$(CC) -O2 -std=c11 -ffp-contract=off -fno-fast-math -fPIC -shared -o $@ $< -lm

clean:
	rm -rf build $(LIB) oracle/_build

.PHONY: all clean
