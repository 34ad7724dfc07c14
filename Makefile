# Build everything in-tree (the .so files travel to the GPU box with the snapshot).
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
# -march=x86-64-v3 mirrors the reference's .cargo/config.toml:23-24; fp-contract off = Rust semantics
CXXFLAGS  := -O3 -std=c++17 -fPIC -march=x86-64-v3 -ffp-contract=off -Wall -Wextra -pthread

PKG := alevin_fry_b200
CSRC := $(PKG)/csrc

all: $(PKG)/libafq.so oracle/libafq_oracle.so synth/libafq_synth.so $(PKG)/libafq_host.so bin/alevin-fry

$(PKG)/libafq.so: $(CSRC)/afq_cuda.cu $(wildcard $(CSRC)/*.cuh) include/afq.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/afq_cuda.cu

oracle/libafq_oracle.so: oracle/afq_oracle.cpp include/afq.h
	$(CXX) $(CXXFLAGS) -shared -o $@ oracle/afq_oracle.cpp

synth/libafq_synth.so: synth/afq_synth.cpp
	$(CXX) $(CXXFLAGS) -shared -o $@ synth/afq_synth.cpp

HOST_SRCS := $(wildcard $(PKG)/host/*.cpp)
HOST_HDRS := $(wildcard $(PKG)/host/*.h) include/afq.h include/afq_host.h

$(PKG)/libafq_host.so: $(HOST_SRCS) $(HOST_HDRS) $(PKG)/libafq.so
	$(CXX) $(CXXFLAGS) -shared -o $@ $(filter-out $(PKG)/host/main.cpp,$(HOST_SRCS)) -L$(PKG) -lafq -lz -Wl,-rpath,'$$ORIGIN'

bin/alevin-fry: $(PKG)/host/main.cpp $(PKG)/libafq_host.so
	mkdir -p bin
	$(CXX) $(CXXFLAGS) -o $@ $(PKG)/host/main.cpp -L$(PKG) -lafq_host -lafq -Wl,-rpath,'$$ORIGIN/../$(PKG)'

clean:
	rm -f $(PKG)/*.so oracle/*.so synth/*.so bin/alevin-fry

.PHONY: all clean
