# Builds libegc_b200.so (sm_100a only) in-tree:  make -j8
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Iinclude --expt-relaxed-constexpr
SRC_DIR   := egc_b200/csrc
BUILD_DIR := build/obj
SRCS      := $(wildcard $(SRC_DIR)/*.cu)
OBJS      := $(patsubst $(SRC_DIR)/%.cu,$(BUILD_DIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(SRC_DIR)/*.cuh) include/egc_b200.h
LIB       := egc_b200/libegc_b200.so

all: $(LIB)

$(BUILD_DIR)/%.o: $(SRC_DIR)/%.cu $(HDRS)
	@mkdir -p $(BUILD_DIR)
	$(NVCC) $(NVCCFLAGS) $(EXTRA) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -lcudart

clean:
	rm -rf build $(LIB)

.PHONY: all clean
