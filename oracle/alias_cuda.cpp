// TEST INFRASTRUCTURE ONLY.  Lets the reference's own CUDA-platform test file
// (platforms/cuda/tests/TestCudaMPIDForce.cpp, compiled unmodified by oracle/Makefile) run against the
// MPIDB200 kernel: the file calls registerMPIDCudaKernelFactories() and then asks for the platform named
// "CUDA" (TestCudaMPIDForce.cpp:60,1002,1709-1711), so this shim registers the MPIDB200 platform under that name.
extern "C" void registerMPIDB200KernelFactoriesAs(const char* platformName);
extern "C" void registerMPIDCudaKernelFactories() { registerMPIDB200KernelFactoriesAs("CUDA"); }
