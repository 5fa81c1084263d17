//! Replaces src/core/kernel.zig (KernelsSet: run-time OpenCL-C JIT + per-queue kernel cache, :123-415).
//! Kernels of the B200 backend are ahead-of-time template instantiations inside libwekua_b200.so, so nothing is
//! compiled or cached at run time; the only declaration other modules still name is the error set.

/// src/core/kernel.zig:9 -- raised by the library as WK_ERR_TYPE_NOT_SUPPORTED (b200.check)
pub const Errors = error{TypeNotSupported};
