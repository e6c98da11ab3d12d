# Builds the C-ABI shared library (sm_100a only) and the CPU oracle.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
CSRC      := border_b200/csrc
OBJDIR    := build/obj
LIB       := border_b200/libborder_b200.so
SRCS      := common.cu replay.cu nn.cu agent.cu dqn.cu sac.cu iqn.cu tc_gemm.cu tma_gemm.cu tma_gemm_fast.cu conv1_tc.cu atari.cu
SRCS      := $(filter $(notdir $(wildcard $(CSRC)/*.cu)),$(SRCS))
OBJS      := $(patsubst %.cu,$(OBJDIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.inc) include/border_b200.h

HOSTLIB   := border_b200/libborder_host.so

all: $(LIB) $(HOSTLIB) oracle

$(OBJDIR)/replay.o: $(CSRC)/replay.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

# host-side loops (Trainer / train_async mirror) in plain C++ over the C ABI only
$(HOSTLIB): border_b200/host/host_capi.cpp border_b200/host/border_host.hpp include/border_host.h include/border_b200.h $(LIB)
	g++ -std=c++17 -O2 -fPIC -shared -Wall -Iinclude border_b200/host/host_capi.cpp -o $@ -Lborder_b200 -lborder_b200 -Wl,-rpath,'$$ORIGIN' -lpthread

oracle: oracle/_build/libreplay_oracle.so oracle/_build/varstore_oracle

# libtorch C++ leg of the oracle (TEST INFRASTRUCTURE): tch's VarStore save / load calls and the real torch::optim::Adam
TORCH_DIR := $(shell python -c "import torch, os; print(os.path.dirname(torch.__file__))" 2>/dev/null)
TORCH_ABI := $(shell python -c "import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))" 2>/dev/null)
oracle/_build/varstore_oracle: oracle/varstore_oracle.cpp
	@mkdir -p oracle/_build
	g++ -std=c++17 -O1 -D_GLIBCXX_USE_CXX11_ABI=$(TORCH_ABI) -I$(TORCH_DIR)/include -I$(TORCH_DIR)/include/torch/csrc/api/include \
	    $< -o $@ -L$(TORCH_DIR)/lib -ltorch -ltorch_cpu -lc10 -Wl,-rpath,$(TORCH_DIR)/lib

oracle/_build/libreplay_oracle.so: oracle/replay_oracle.c
	@mkdir -p oracle/_build
	gcc -O2 -fPIC -shared -o $@ $< -lm

clean:
	rm -rf build $(LIB) $(HOSTLIB) oracle/_build

.PHONY: all oracle clean
