# Builds the C-ABI library in-tree (sm_100a only) and the nothing-else. `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
CSRC := dir_b200/csrc
SRCS := $(CSRC)/conv_simt.cu $(CSRC)/conv_tc.cu $(CSRC)/conv_tf32.cu $(CSRC)/conv_halo.cu $(CSRC)/conv_b2b.cu $(CSRC)/stem_pool.cu $(CSRC)/elementwise.cu $(CSRC)/joint.cu $(CSRC)/ste_tc.cu $(CSRC)/gcn_tc.cu $(CSRC)/fusion.cu $(CSRC)/mano.cu $(CSRC)/eval_metric.cu $(CSRC)/engine.cu $(CSRC)/capi.cu
OBJS := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
HDRS := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr

all: dir_b200/libdirb200.so

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

dir_b200/libdirb200.so: $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -lcuda -ldl

clean:
	rm -rf build dir_b200/libdirb200.so
