# Builds the C-ABI library (sm_100a only) and the C oracle.
NVCC ?= nvcc
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -diag-suppress 177 -Wno-deprecated-gpu-targets
CSRC = $(wildcard flowcompare_b200/csrc/*.cu)
OBJS = $(patsubst flowcompare_b200/csrc/%.cu,build/%.o,$(CSRC))
LIB = flowcompare_b200/libflowcompare_b200.so

all: $(LIB) oracle

build/%.o: flowcompare_b200/csrc/%.cu $(wildcard flowcompare_b200/csrc/*.cuh) include/flowcompare_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -lcuda

oracle: oracle/_build/libfc_oracle.so
oracle/_build/libfc_oracle.so: $(wildcard oracle/*.c)
	@mkdir -p oracle/_build
	gcc -O2 -ffp-contract=off -fPIC -shared -o $@ $^ -lm

clean:
	rm -rf build $(LIB) oracle/_build
.PHONY: all oracle clean
