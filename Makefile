# Builds libpaimon_gpu.so (sm_100a only) in-tree, and the parity oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude -Ipaimon_b200/csrc
SRCS := paimon_b200/csrc/merge.cu paimon_b200/csrc/emit.cu paimon_b200/csrc/api.cu \
	paimon_b200/csrc/parquet_decode.cu paimon_b200/csrc/parquet_encode.cu paimon_b200/csrc/parquet_meta.cc \
	paimon_b200/csrc/arrow_export.cu paimon_b200/csrc/upload.cu paimon_b200/csrc/readback.cu paimon_b200/csrc/orc_decode.cu paimon_b200/csrc/orc_meta.cc
HDRS := include/paimon_gpu.h paimon_b200/csrc/pg_internal.h paimon_b200/csrc/device_utils.cuh paimon_b200/csrc/parquet_meta.h \
	paimon_b200/csrc/zstd_device.cuh paimon_b200/csrc/inflate_device.cuh paimon_b200/csrc/orc_device.cuh paimon_b200/csrc/orc_meta.h \
	paimon_b200/csrc/scan_kernels.cuh
LIB := paimon_b200/libpaimon_gpu.so

all: $(LIB) oracle

# one object per source (build/ is scratch), linked into the shared library
OBJDIR ?= build
OBJS := $(patsubst paimon_b200/csrc/%,$(OBJDIR)/%.o,$(SRCS))

$(OBJDIR)/%.o: paimon_b200/csrc/% $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) $(EXTRA_DEFS) -x cu -c -o $@ $<

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

ptxas-info: $(SRCS) $(HDRS)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c -o /dev/null paimon_b200/csrc/merge.cu

# the JNI shim against the JNI specification's signatures (no JDK in the image: jni/stub/jni.h)
jni-check:
	g++ -std=c++17 -fsyntax-only -Wall -Ijni/stub -Iinclude jni/paimon_gpu_jni.cc

oracle:
	$(MAKE) -C oracle -s

clean:
	rm -rf $(LIB) $(OBJDIR)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean ptxas-info jni-check
