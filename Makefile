# Builds libmfsgd.so (sm_100a only), the CPU oracle (test infrastructure) and the C ABI harness.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v
PKG       := matrixfactorizationsgd.java_b200
CSRC      := $(PKG)/csrc
OBJDIR    := build/obj
LIB       := $(PKG)/lib/libmfsgd.so
OBJS      := $(OBJDIR)/engine.o $(OBJDIR)/kernels_update.o $(OBJDIR)/kernels_hot.o $(OBJDIR)/kernels_layout.o $(OBJDIR)/kernels_eval.o $(OBJDIR)/kernels_diag.o $(OBJDIR)/kernels_ring.o $(OBJDIR)/ratings_io.o

all: $(LIB) oracle harness host tools/l2_peak

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/kernels.cuh $(CSRC)/update_math.cuh $(CSRC)/run_plan.hpp include/mfsgd.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

$(OBJDIR)/ratings_io.o: $(CSRC)/ratings_io.cpp include/mfsgd.h
	@mkdir -p $(OBJDIR)
	g++ -O2 -std=c++17 -Wall -Wextra -fPIC -fvisibility=hidden -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(PKG)/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl -lpthread

oracle:
	$(MAKE) -s -C oracle

harness: tests/c/abi_harness
tests/c/abi_harness: tests/c/abi_harness.c include/mfsgd.h
	gcc -O1 -Wall -Wextra -Iinclude -o $@ $< -ldl

host: host/factorize_demo
host/factorize_demo: host/factorize_demo.cpp host/MatrixFactorizationSGD.hpp include/mfsgd.h
	g++ -O1 -std=c++17 -Wall -Wextra -o $@ $< -ldl

# measured ceilings of the update path's access pattern (bench/profile aid, not product code)
tools/l2_peak: tools/l2_peak.cu
	$(NVCC) $(ARCH) -O3 -lineinfo -o $@ $<

clean:
	rm -rf build $(LIB) tests/c/abi_harness host/factorize_demo tools/l2_peak
	$(MAKE) -s -C oracle clean

# The Java host and the reference stand-in (JDK 22+; this image has none, so the target only reports that)
java:
	@if command -v javac >/dev/null 2>&1; then \
	    mkdir -p build/java && javac --release 22 -d build/java java/MatrixFactorizationSGDGpu.java baseline/java/MatrixFactorizationSGD.java && \
	    echo "built build/java; run: java --enable-native-access=ALL-UNNAMED -cp build/java -Dmfsgd.lib=$(abspath $(LIB)) MatrixFactorizationSGDGpu"; \
	else echo "no JDK on PATH: java/ and baseline/java/ stay source-only (see java/README.md)"; fi

.PHONY: all oracle harness host java clean
