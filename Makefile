# Builds the C-ABI library (sm_100a only) and the CPU oracle.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v
SRC := $(wildcard pienerf_b200/csrc/*.cu)
OBJ := $(patsubst pienerf_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := pienerf_b200/lib/libpienerf_b200.so

all: $(LIB) oracle

build/%.o: pienerf_b200/csrc/%.cu $(wildcard pienerf_b200/csrc/*.cuh) include/pienerf_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p pienerf_b200/lib
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ) -lcudart

oracle: oracle/_build/libsim_oracle.so
oracle/_build/libsim_oracle.so: oracle/sim_oracle.c
	@mkdir -p oracle/_build
	gcc -O2 -fopenmp -shared -fPIC -o $@ $< -lm

# developer A/B builds: make variant TAG=x EXTRA="-DPN_FOO=1"  ->  pienerf_b200/lib/libpienerf_b200_x.so (load with PN_LIB=...)
variant:
	@mkdir -p build_$(TAG) pienerf_b200/lib
	@for f in $(SRC); do b=$$(basename $$f .cu); echo "[$(TAG)] $$b"; $(NVCC) $(NVFLAGS) $(EXTRA) -c $$f -o build_$(TAG)/$$b.o 2> build_$(TAG)/$$b.ptxas.log || { cat build_$(TAG)/$$b.ptxas.log; exit 1; }; done
	$(NVCC) -shared $(ARCH) -o pienerf_b200/lib/libpienerf_b200_$(TAG).so build_$(TAG)/*.o -lcudart

clean:
	rm -rf build $(LIB) oracle/_build

.PHONY: all oracle clean variant
