# Top-level build: libqcs.so (CUDA, sm_100a), the C host driver, the oracles.
#
#   make            everything
#   make lib        quantumcomputer_b200/lib/libqcs.so
#   make host       quantumcomputer_b200/bin/qc_shor_b200
#   make oracle     oracle/_build + oracle/_ref (test infrastructure)
#   make dropin     oracle/_ref/qc_shor_{ref,dropin,dropin_fused}: the reference program with INTEGRATION.md's edits,
#                   linked against libqcs.so (test infrastructure; only where /root/reference exists)

NVCC      ?= nvcc
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall \
             --fmad=true -Xptxas -v
PKG       := quantumcomputer_b200
CSRC      := $(PKG)/csrc
OBJDIR    := build/obj
LIB       := $(PKG)/lib/libqcs.so
HOSTBIN   := $(PKG)/bin/qc_shor_b200
HOSTLIB   := $(PKG)/lib/libqcshost.so

CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CU_OBJS   := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU_SRCS))

all: lib host oracle dropin

lib: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CSRC)/qcs_internal.h $(CSRC)/qft_common.cuh include/qcs.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIB): $(CU_OBJS)
	@mkdir -p $(PKG)/lib
	$(NVCC) $(ARCH) -shared -o $@ $(CU_OBJS) -ldl

# development build with the pipeline's role timing compiled in (QCS_LIB_PATH=build/timing/libqcs.so QCS_PIPE_TIMING=1)
timing:
	@mkdir -p build/timing
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --fmad=true -DQCS_PIPE_TIMING -shared -o build/timing/libqcs.so $(CU_SRCS) -ldl

host: $(HOSTBIN) $(HOSTLIB)

# the classical half as a shared object, so tests can drive it through ctypes
$(HOSTLIB): $(PKG)/host/mt19937.c $(PKG)/host/mt19937.h $(PKG)/host/shor_classical.c $(PKG)/host/shor_classical.h $(PKG)/host/state_debug.c $(PKG)/host/state_debug.h $(LIB)
	$(CC) -O2 -Wall -fPIC -shared -Iinclude -o $@ $(PKG)/host/mt19937.c $(PKG)/host/shor_classical.c $(PKG)/host/state_debug.c \
	      -L$(PKG)/lib -lqcs -Wl,-rpath,'$$ORIGIN' -lm

$(HOSTBIN): $(PKG)/host/qc_shor_b200.c $(PKG)/host/mt19937.c $(PKG)/host/mt19937.h $(PKG)/host/shor_classical.c $(PKG)/host/shor_classical.h $(LIB)
	@mkdir -p $(PKG)/bin
	$(CC) -O2 -Wall -Iinclude -o $@ $(PKG)/host/qc_shor_b200.c $(PKG)/host/mt19937.c $(PKG)/host/shor_classical.c \
	      -L$(PKG)/lib -lqcs -Wl,-rpath,'$$ORIGIN/../lib' -lm

oracle:
	$(MAKE) -C oracle --no-print-directory

# test infrastructure: never fatal for the product build (tests/test_reference_dropin.py skips without the binaries)
dropin: $(LIB)
	-python oracle/make_dropin.py

clean:
	rm -rf build $(PKG)/lib $(PKG)/bin
	$(MAKE) -C oracle clean

.PHONY: all lib host oracle dropin clean timing
