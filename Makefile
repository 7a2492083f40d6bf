# Build libtethys.so (sm_100a only) and the standalone GPU self-tests. nvcc cross-compiles without a GPU.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function \
             --expt-relaxed-constexpr -cudart static
CSRC      := tethys_speech_b200/csrc
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) include/tethys.h
LIB       := tethys_speech_b200/libtethys.so

all: $(LIB) tools/selftest_gemm tools/selftest_attn tools/nvml_sampler

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(OBJS) -ldl

tools/selftest_gemm: tools/selftest_gemm.cu $(LIB) include/tethys.h
	$(NVCC) $(ARCH) -O2 -std=c++17 -cudart static -o $@ tools/selftest_gemm.cu -Ltethys_speech_b200 -ltethys \
	    -Xlinker -rpath -Xlinker '$$ORIGIN/../tethys_speech_b200'

tools/selftest_attn: tools/selftest_attn.cu $(LIB) include/tethys.h
	$(NVCC) $(ARCH) -O2 -std=c++17 -cudart static -o $@ tools/selftest_attn.cu -Ltethys_speech_b200 -ltethys \
	    -Xlinker -rpath -Xlinker '$$ORIGIN/../tethys_speech_b200'

tools/nvml_sampler: tools/nvml_sampler.cpp
	g++ -O2 -std=c++17 -Wall -o $@ tools/nvml_sampler.cpp -ldl

clean:
	rm -rf build $(LIB) tools/selftest_gemm tools/selftest_attn tools/nvml_sampler

.PHONY: all clean
