# Builds everything in-tree (the .so files travel to the GPU box with the snapshot).
#   ipu_ray_lib_b200/libb200rt.so        CUDA trace path behind include/b200rt.h   (sm_100a only)
#   ipu_ray_lib_b200/libb200rt_scene.so  host-side scene utilities, include/b200rt_scene.h
#   ipu_ray_lib_b200/trace               CLI with the reference's flags (host/trace_main.cpp)
#   oracle/liboracle_port.so, oracle/_ref/liboracle_ref.so   CPU checkers (tests / bench baseline only)
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      := /usr/bin/g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
# Bit-exact parity with the FMA-free reference CPU path: no contraction, IEEE div/sqrt, no FTZ.
NVFLAGS  := $(ARCH) -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -ccbin $(CXX) \
            --fmad=false -prec-div=true -prec-sqrt=true -ftz=false -diag-suppress 549
# The NIF MLP is tolerance-checked, not bit-compared: contraction allowed there.
NVFLAGS_NIF := $(ARCH) -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -ccbin $(CXX)
HOSTFLAGS := -std=c++17 -O2 -fPIC -ffp-contract=off -fno-fast-math -Wall -Wextra

PKG   := ipu_ray_lib_b200
CSRC  := $(PKG)/csrc
HOST  := $(PKG)/host

all: $(PKG)/libb200rt.so $(PKG)/libb200rt_scene.so $(PKG)/trace oracle tests/libhostpair.so

$(CSRC)/b200rt.o: $(CSRC)/b200rt.cu $(CSRC)/trace_kernels.cuh $(CSRC)/wavefront.cuh $(CSRC)/rt_device.cuh $(CSRC)/rt_prims.h $(CSRC)/pair_build.hpp $(CSRC)/scene_tables.hpp $(CSRC)/rt_math.h \
                  $(CSRC)/sin_deg_table.inc $(CSRC)/nif.cuh include/b200rt.h
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(CSRC)/nif.o: $(CSRC)/nif.cu $(CSRC)/nif.cuh $(CSRC)/nif_tc.cuh include/b200rt.h
	$(NVCC) $(NVFLAGS_NIF) -c -o $@ $<

$(PKG)/libb200rt.so: $(CSRC)/b200rt.o $(CSRC)/nif.o
	$(NVCC) $(ARCH) -shared -ccbin $(CXX) -o $@ $^ -cudart static

HOST_SRCS := $(HOST)/scene_build.cpp $(HOST)/gltf_import.cpp $(HOST)/image_io.cpp $(HOST)/scene_capi.cpp \
             $(HOST)/scene_import.cpp $(HOST)/keras_hdf5.cpp
$(PKG)/libb200rt_scene.so: $(HOST_SRCS) $(HOST)/scene_build.hpp $(HOST)/rt_types.hpp $(HOST)/mini_json.hpp \
                           $(CSRC)/rt_math.h include/b200rt_scene.h include/b200rt.h
	$(CXX) $(HOSTFLAGS) -shared -o $@ $(HOST_SRCS)

$(PKG)/trace: $(HOST)/trace_main.cpp $(HOST)/B200Scene.hpp $(PKG)/libb200rt.so $(PKG)/libb200rt_scene.so
	$(CXX) $(HOSTFLAGS) -o $@ $(HOST)/trace_main.cpp -L$(PKG) -lb200rt -lb200rt_scene -lpthread \
	    -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle all

# test infrastructure: the traversal core of the kernels (csrc/rt_prims.h) compiled for the host, checked against the oracle
tests/libhostpair.so: tests/host_pair_check.cpp $(CSRC)/rt_prims.h $(CSRC)/pair_build.hpp $(CSRC)/scene_tables.hpp $(CSRC)/rt_math.h \
                      $(CSRC)/sin_deg_table.inc include/b200rt.h
	$(CXX) $(HOSTFLAGS) -I/usr/local/cuda/include -shared -o $@ $<

# measured-slower kernel variants, never part of the product library (scripts/experiments/README.md):
#   variants/libb200rt_pair.so = the CTA-pair NIF kernel (csrc/nif_tc_pair2.cuh), run with B200RT_LIB=... B200RT_NIF_PAIR=2
experiments: $(CSRC)/b200rt.o
	mkdir -p $(PKG)/variants
	$(NVCC) $(NVFLAGS_NIF) -DB200RT_NIF_PAIR_KERNEL -c -o $(PKG)/variants/nif_pair.o $(CSRC)/nif.cu
	$(NVCC) $(ARCH) -shared -ccbin $(CXX) -o $(PKG)/variants/libb200rt_pair.so $(CSRC)/b200rt.o $(PKG)/variants/nif_pair.o -cudart static

sass: $(PKG)/libb200rt.so
	/usr/local/cuda/bin/cuobjdump -sass $(PKG)/libb200rt.so > /tmp/b200rt.sass

clean:
	rm -f $(CSRC)/*.o $(PKG)/*.so $(PKG)/trace
	$(MAKE) -C oracle clean

.PHONY: all oracle clean sass experiments
