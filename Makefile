# Builds sopht_b200/lib/libsopht_b200.so (sm_100a only) and the oracle's C restatement.
NVCC      ?= /usr/local/cuda/bin/nvcc
CUDA_HOME ?= /usr/local/cuda
ARCH      := -gencode arch=compute_100a,code=sm_100a
# EXTRA / BUILD_DIR / LIBNAME: experiment builds next to the product library, e.g.
#   make lib EXTRA=-DSOPHT_FFT_1024_E16 BUILD_DIR=build_e16 LIBNAME=libsopht_b200_e16.so
# selected at run time with SOPHT_B200_LIB=<path> (sopht_b200/_lib.py)
EXTRA     ?=
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude -Isopht_b200/csrc \
             --expt-relaxed-constexpr -Xptxas -warn-spills $(EXTRA)
SRC_DIR   := sopht_b200/csrc
BUILD_DIR ?= build
LIB_DIR   := sopht_b200/lib
LIBNAME   ?= libsopht_b200.so
LIB       := $(LIB_DIR)/$(LIBNAME)

CU_SRCS := $(wildcard $(SRC_DIR)/*.cu)
OBJS    := $(patsubst $(SRC_DIR)/%.cu,$(BUILD_DIR)/%.o,$(CU_SRCS))
HDRS    := $(wildcard $(SRC_DIR)/*.cuh) include/sopht_b200.h

all: $(LIB) oracle

lib: $(LIB)

$(BUILD_DIR)/%.o: $(SRC_DIR)/%.cu $(HDRS)
	@mkdir -p $(BUILD_DIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(LIB_DIR)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -L$(CUDA_HOME)/lib64 -lcufft -Xlinker -rpath,$(CUDA_HOME)/lib64

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build build_* $(LIB_DIR) oracle/_build

.PHONY: all lib clean oracle
