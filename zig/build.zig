//! build.zig of the B200 overlay: the reference's module graph (build.zig:13-80) with the `opencl` dependency replaced by
//! the in-tree stub and every module linked against libwekua_b200.so.
//!
//!   zig build -Dwekua-b200=/path/to/this/repo            # -> links <repo>/wekua_b200/libwekua_b200.so
//!
//! Files not present in this overlay are taken from the reference tree unchanged (see README.md).
const std = @import("std");

pub fn build(b: *std.Build) void {
    const target = b.standardTargetOptions(.{});
    const optimize = b.standardOptimizeOption(.{});
    const repo = b.option([]const u8, "wekua-b200", "root of the wekua_b200 repository (holds wekua_b200/libwekua_b200.so)") orelse "..";

    const opencl_module = b.addModule("opencl", .{
        .root_source_file = b.path("src/opencl_stub/opencl.zig"),
        .target = target,
        .optimize = optimize,
    });

    const utils_module = b.addModule("utils", .{
        .root_source_file = b.path("src/utils/utils.zig"),
        .target = target,
        .optimize = optimize,
    });

    const core_module = b.addModule("core", .{
        .root_source_file = b.path("src/core/main.zig"),
        .target = target,
        .optimize = optimize,
        .link_libc = true,
    });
    core_module.addImport("opencl", opencl_module);
    // the one native dependency: the C-ABI library (CUDA runtime statically linked inside it)
    core_module.addLibraryPath(.{ .cwd_relative = b.pathJoin(&.{ repo, "wekua_b200" }) });
    core_module.addRPath(.{ .cwd_relative = b.pathJoin(&.{ repo, "wekua_b200" }) });
    core_module.linkSystemLibrary("wekua_b200", .{});

    const tensor_module = b.addModule("tensor", .{
        .root_source_file = b.path("src/tensor/main.zig"),
        .target = target,
        .optimize = optimize,
    });
    tensor_module.addImport("opencl", opencl_module);
    tensor_module.addImport("core", core_module);
    tensor_module.addImport("utils", utils_module);

    const blas_module = b.addModule("blas", .{
        .root_source_file = b.path("src/blas/main.zig"),
        .target = target,
        .optimize = optimize,
    });
    blas_module.addImport("opencl", opencl_module);
    blas_module.addImport("core", core_module);
    blas_module.addImport("utils", utils_module);
    blas_module.addImport("tensor", tensor_module);

    const math_module = b.addModule("math", .{
        .root_source_file = b.path("src/math/main.zig"),
        .target = target,
        .optimize = optimize,
    });
    math_module.addImport("opencl", opencl_module);
    math_module.addImport("core", core_module);
    math_module.addImport("utils", utils_module);
    math_module.addImport("tensor", tensor_module);

    const nn_module = b.addModule("nn", .{
        .root_source_file = b.path("src/nn/main.zig"),
        .target = target,
        .optimize = optimize,
    });
    nn_module.addImport("opencl", opencl_module);
    nn_module.addImport("core", core_module);
    nn_module.addImport("utils", utils_module);
    nn_module.addImport("tensor", tensor_module);
    nn_module.addImport("blas", blas_module);
    nn_module.addImport("math", math_module);

    const wekua_module = b.addModule("wekua", .{
        .root_source_file = b.path("src/wekua.zig"),
        .target = target,
        .optimize = optimize,
    });
    wekua_module.addImport("opencl", opencl_module);
    wekua_module.addImport("core", core_module);
    wekua_module.addImport("tensor", tensor_module);
    wekua_module.addImport("utils", utils_module);
    wekua_module.addImport("blas", blas_module);
    wekua_module.addImport("math", math_module);
    wekua_module.addImport("nn", nn_module);

    // examples/xor_neural_network.zig, unchanged (zig build run_example -Dexample=xor_neural_network)
    const example_name = b.option([]const u8, "example", "name of the example under examples/") orelse "xor_neural_network";
    const example = b.addExecutable(.{
        .name = example_name,
        .root_module = b.createModule(.{
            .root_source_file = b.path(b.fmt("examples/{s}.zig", .{example_name})),
            .target = target,
            .optimize = optimize,
        }),
    });
    example.root_module.addImport("wekua", wekua_module);
    const run_example = b.addRunArtifact(example);
    b.step("run_example", "Run an example on the B200 backend").dependOn(&run_example.step);

    const test_step = b.step("test", "Run unit tests");
    inline for (.{ .{ core_module, "core" }, .{ tensor_module, "tensor" }, .{ blas_module, "blas" }, .{ math_module, "math" }, .{ nn_module, "nn" } }) |mt| {
        const t = b.addTest(.{ .root_module = mt[0], .use_llvm = true, .name = mt[1] });
        const run = b.addRunArtifact(t);
        run.has_side_effects = true;
        test_step.dependOn(&run.step);
    }
}
