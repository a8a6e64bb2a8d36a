// skrample_b200 - the plan-cache hit path of the Python layer, without the Python.
//
// `sampler.sample()` on a step that was taken before comes down to: check that the tensors are what the plan was
// made for, allocate the outputs, read the current stream, call skr_plan_launch.  In Python that is ~40 attribute reads,
// three torch.empty calls and a ctypes call - 25 us on the benchmark box against a 6 us kernel.  This module does the
// same with the ATen C++ API in one call (~8 us).  It is glue between torch and the C ABI (include/skrample_b200.h):
// no arithmetic, no kernels, no state beyond the address of skr_plan_launch handed over by native.py.  When the
// module is absent native.py runs the identical sequence in Python.
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <cstdint>
#include <vector>

namespace {

using plan_launch_fn = int (*)(const void* plan, const void* const* tensors, int64_t numel, const void* draws, void* stream);
plan_launch_fn g_plan_launch = nullptr;

constexpr int kMaxTensors = 40;  // SKR_MAX_INPUTS + SKR_MAX_OUTPUTS

inline int code_of(at::ScalarType t) {
    switch (t) {
        case at::kFloat: return 0;
        case at::kDouble: return 1;
        case at::kBFloat16: return 2;
        case at::kHalf: return 3;
        default: return -1;
    }
}
inline at::ScalarType type_of(int code) {
    switch (code) {
        case 1: return at::kDouble;
        case 2: return at::kBFloat16;
        case 3: return at::kHalf;
        default: return at::kFloat;
    }
}

void bind(uintptr_t plan_launch) { g_plan_launch = reinterpret_cast<plan_launch_fn>(plan_launch); }

// plans:  dict { signature (int: 2 bits of dtype code per input, first input lowest) -> (plan handle, bytes of output
//         dtype codes) } of one compiled program;
// inputs: list of tensors in program order;  draws: address of the skr_philox tables (0: none).
// Returns the list of output tensors;  None when the tensors do not qualify (not CUDA, mixed devices / shapes,
// non-contiguous, unsupported dtype);  the signature (int) when `plans` has no entry for it (the caller creates the
// plan and calls again);  a negative or CUDA status wrapped in a 1-tuple when the launch failed.
py::object launch(py::dict plans, py::list inputs, uintptr_t draws, bool account) {
    const Py_ssize_t n = PyList_GET_SIZE(inputs.ptr());
    if (n < 1 || n > 32 || !g_plan_launch) return py::none();
    const void* table[kMaxTensors];
    at::Tensor first;
    uint64_t signature = 0;
    c10::DeviceIndex device = -1;
    int64_t bytes = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* item = PyList_GET_ITEM(inputs.ptr(), i);
        if (!THPVariable_Check(item)) return py::none();
        const at::Tensor& t = THPVariable_Unpack(item);
        if (!t.is_cuda() || !t.is_contiguous()) return py::none();
        const int code = code_of(t.scalar_type());
        if (code < 0) return py::none();
        if (i == 0) {
            first = t;
            device = t.get_device();
        } else if (t.get_device() != device || !t.sizes().equals(first.sizes())) {
            return py::none();
        }
        signature |= (uint64_t)code << (2 * i);
        table[i] = t.data_ptr();
        if (account) bytes += t.numel() * (int64_t)t.element_size();
    }
    PyObject* key = PyLong_FromUnsignedLongLong(signature);
    PyObject* entry = PyDict_GetItem(plans.ptr(), key);  // borrowed
    if (!entry) return py::reinterpret_steal<py::object>(key);
    Py_DECREF(key);
    const uintptr_t plan = PyLong_AsUnsignedLongLong(PyTuple_GET_ITEM(entry, 0));
    PyObject* out_codes = PyTuple_GET_ITEM(entry, 1);
    const Py_ssize_t n_out = PyBytes_GET_SIZE(out_codes);
    const char* codes = PyBytes_AS_STRING(out_codes);
    if (n + n_out > kMaxTensors) return py::none();

    c10::cuda::OptionalCUDAGuard guard;
    if (c10::cuda::current_device() != device) guard.set_index(device);
    py::list outputs(n_out);
    const auto options = first.options();
    for (Py_ssize_t j = 0; j < n_out; ++j) {
        at::Tensor out = at::empty(first.sizes(), options.dtype(type_of(codes[j])));
        table[n + j] = out.data_ptr();
        if (account) bytes += out.numel() * (int64_t)out.element_size();
        PyList_SET_ITEM(outputs.ptr(), j, THPVariable_Wrap(std::move(out)));
    }
    void* stream = c10::cuda::getCurrentCUDAStream(device).stream();
    const int status = g_plan_launch(reinterpret_cast<const void*>(plan), table, first.numel(), reinterpret_cast<const void*>(draws), stream);
    if (status != 0) return py::make_tuple(status);
    if (account) return py::make_tuple(outputs, bytes);
    return outputs;
}

}  // namespace

PYBIND11_MODULE(_fast, m) {
    m.doc() = "skrample_b200: plan-cache hit path (torch tensors -> skr_plan_launch) in one call";
    m.def("bind", &bind, "hand over the address of skr_plan_launch");
    m.def("launch", &launch, py::arg("plans"), py::arg("inputs"), py::arg("draws") = 0, py::arg("account") = false);
}
