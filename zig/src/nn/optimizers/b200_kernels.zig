//! The device side of src/nn/optimizers/{gdm,adagrad,rmsprop}.zig on the B200 backend (+ Adam, whose Zig file is empty in
//! the reference).  Each of those files is host glue (state tensors, the Optimizer vtable, the walk over cache.slots)
//! around ONE private function that launches a kernel; the overlay replaces the body of that function:
//!
//!     gdm.zig:111      fn executeGDM(self, pipeline, x, gradient, velocity)            -> kernels.gdm(T, pipeline, x, gradient, velocity, self.config.lr, self.config.beta)
//!     adagrad.zig:110  fn executeAdagrad(self, pipeline, x, gradient, gradient_history) -> kernels.adagrad(T, pipeline, x, gradient, gradient_history, self.config.lr)
//!     rmsprop.zig:111  fn executeRMSProp(self, pipeline, x, gradient, gradient_history) -> kernels.rmsprop(T, pipeline, x, gradient, gradient_history, self.config.lr, self.config.gamma)
//!
//! Scalars travel by pointer to ONE element of T, which also removes the reference's sizeof(T)-vs-vector argument mismatch
//! (SURVEY Q4).  gd.zig needs no change: it calls blas.axpy.  `stepMulti` is the one-launch form of a whole `step`.
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("tensor");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

/// gdm.cl:3-33: v = beta v + lr g; x -= v
pub fn gdm(comptime T: type, pipeline: *Pipeline, x: *Tensor(T), gradient: *Tensor(T), velocity: *Tensor(T), lr: T, beta: T) TensorErrors!void {
    try b200.check(b200.wk_gdm(pipeline.q(), core.types.getTypeIndex(T), x.buffer, gradient.buffer, velocity.buffer, @ptrCast(&lr), @ptrCast(&beta), x.dimensions.number_of_elements));
}

/// adagrad.cl:3-48: h += g^2; x -= lr g / (sqrt(h) + FLT_EPSILON)
pub fn adagrad(comptime T: type, pipeline: *Pipeline, x: *Tensor(T), gradient: *Tensor(T), gradient_history: *Tensor(T), lr: T) TensorErrors!void {
    try b200.check(b200.wk_adagrad(pipeline.q(), core.types.getTypeIndex(T), x.buffer, gradient.buffer, gradient_history.buffer, @ptrCast(&lr), x.dimensions.number_of_elements));
}

/// rmsprop.cl:3-57: h = gamma h + (1 - gamma) g^2; x -= lr g / (sqrt(h) + FLT_EPSILON)
pub fn rmsprop(comptime T: type, pipeline: *Pipeline, x: *Tensor(T), gradient: *Tensor(T), gradient_history: *Tensor(T), lr: T, gamma: T) TensorErrors!void {
    try b200.check(b200.wk_rmsprop(pipeline.q(), core.types.getTypeIndex(T), x.buffer, gradient.buffer, gradient_history.buffer, @ptrCast(&lr), @ptrCast(&gamma), x.dimensions.number_of_elements));
}

/// Adam with bias correction by the step count t (new op: src/nn/optimizers/adam.zig is a 0-byte file)
pub fn adam(comptime T: type, pipeline: *Pipeline, x: *Tensor(T), gradient: *Tensor(T), m: *Tensor(T), v: *Tensor(T), lr: T, beta1: T, beta2: T, eps: T, t: u64) TensorErrors!void {
    try b200.check(b200.wk_adam(pipeline.q(), core.types.getTypeIndex(T), x.buffer, gradient.buffer, m.buffer, v.buffer, @ptrCast(&lr), @ptrCast(&beta1), @ptrCast(&beta2), @ptrCast(&eps), t, x.dimensions.number_of_elements));
}

/// One launch for a whole Optimizer.step: `params` lists (x, gradient, state0, state1, n) of every weight and bias tensor the
/// reference's step loops visit one kernel at a time (gd.zig:55-94, rmsprop.zig:168-202); results are bit-identical.
/// kind = b200.OPT_*; h0..h2 = beta | gamma | (beta1, beta2, eps).
pub fn stepMulti(comptime T: type, pipeline: *Pipeline, kind: i32, params: []const b200.OptParam, lr: T, h0: ?T, h1: ?T, h2: ?T, t: u64) TensorErrors!void {
    if (params.len == 0) return;
    try b200.check(b200.wk_optimizer_step_multi(pipeline.q(), core.types.getTypeIndex(T), kind, params.ptr, @intCast(params.len), @ptrCast(&lr), b200.optPtr(T, &h0), b200.optPtr(T, &h1), b200.optPtr(T, &h2), t));
}
