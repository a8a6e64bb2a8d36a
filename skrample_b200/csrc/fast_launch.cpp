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
#include <torch/csrc/Dtype.h>

#include "../../include/skrample_b200.h"

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

// ---- the whole plan hit: bind the step's tensors by role, fill the Philox key tables, launch ---------------------
//
// roles: tuple of role tuples as written by sampling/plan.py: (0,) sample, (1,) prediction, (2,) noise,
// (4, back, field[, key]) field 0/1/2 = sample / prediction / noise of previous[-back], 3 = its x-hat cache (valid when
// the cache's key equals `key`), (9, inner) = the tensor of the lazy draw that role `inner` names (materialised on
// demand).  The first n_inputs roles are tensors, the rest lazy draws (PhiloxDraw: .seeds, .streams, .item_numel,
// .numel, .dtype).  Returns what launch() returns, or None when a role cannot be resolved.

struct Names {
    PyObject *sample, *prediction, *noise, *dict, *xhat, *materialize, *seeds, *streams, *item_numel, *numel, *dtype, *tensor, *offset_inner, *offset_scale;
    Names()
        : sample(PyUnicode_InternFromString("sample")), prediction(PyUnicode_InternFromString("prediction")),
          noise(PyUnicode_InternFromString("noise")), dict(PyUnicode_InternFromString("__dict__")),
          xhat(PyUnicode_InternFromString("_skr_xhat")), materialize(PyUnicode_InternFromString("materialize")),
          seeds(PyUnicode_InternFromString("seeds")), streams(PyUnicode_InternFromString("streams")),
          item_numel(PyUnicode_InternFromString("item_numel")), numel(PyUnicode_InternFromString("numel")),
          dtype(PyUnicode_InternFromString("dtype")), tensor(PyUnicode_InternFromString("_tensor")),
          offset_inner(PyUnicode_InternFromString("offset_inner")), offset_scale(PyUnicode_InternFromString("offset_scale")) {}
};
const Names& names() {
    static const Names n;
    return n;
}

// new reference, or nullptr (no Python error left set) when the role does not resolve
PyObject* resolve(PyObject* role, PyObject* packed, PyObject* previous) {
    const Names& n = names();
    const long kind = PyLong_AsLong(PyTuple_GET_ITEM(role, 0));
    if (kind == 0) return PyObject_GetAttr(packed, n.sample);
    if (kind == 1) return PyObject_GetAttr(packed, n.prediction);
    if (kind == 2) return PyObject_GetAttr(packed, n.noise);
    if (kind == 4) {
        const Py_ssize_t back = PyLong_AsSsize_t(PyTuple_GET_ITEM(role, 1));
        const Py_ssize_t count = PySequence_Size(previous);
        if (back < 1 || back > count) return nullptr;
        PyObject* entry = PySequence_GetItem(previous, count - back);  // new reference
        if (!entry) { PyErr_Clear(); return nullptr; }
        const long field = PyLong_AsLong(PyTuple_GET_ITEM(role, 2));
        PyObject* value = nullptr;
        if (field == 3) {
            PyObject* dict = PyObject_GetAttr(entry, n.dict);
            PyObject* held = dict ? PyDict_GetItem(dict, n.xhat) : nullptr;  // borrowed
            if (held && PyTuple_Check(held) && PyTuple_GET_SIZE(held) == 2 && PyTuple_GET_SIZE(role) == 4 &&
                PyObject_RichCompareBool(PyTuple_GET_ITEM(held, 0), PyTuple_GET_ITEM(role, 3), Py_EQ) == 1) {
                value = PyTuple_GET_ITEM(held, 1);
                Py_INCREF(value);
            }
            Py_XDECREF(dict);
        } else {
            value = PyObject_GetAttr(entry, field == 0 ? n.sample : field == 1 ? n.prediction : n.noise);
        }
        Py_DECREF(entry);
        if (!value) PyErr_Clear();
        return value;
    }
    if (kind == 9) {
        PyObject* draw = resolve(PyTuple_GET_ITEM(role, 1), packed, previous);
        if (!draw) return nullptr;
        PyObject* tensor = PyObject_HasAttr(draw, n.materialize) ? PyObject_CallMethodNoArgs(draw, n.materialize) : nullptr;
        Py_DECREF(draw);
        if (!tensor) PyErr_Clear();
        return tensor;
    }
    return nullptr;
}

bool fill_draw(skr_philox& table, PyObject* draw, int64_t* numel) {
    const Names& n = names();
    PyObject* seeds = PyObject_GetAttr(draw, n.seeds);
    PyObject* streams = PyObject_GetAttr(draw, n.streams);
    PyObject* item_numel = PyObject_GetAttr(draw, n.item_numel);
    PyObject* total = PyObject_GetAttr(draw, n.numel);
    PyObject* dtype = PyObject_GetAttr(draw, n.dtype);
    PyObject* inner = PyObject_GetAttr(draw, n.offset_inner);
    PyObject* scale = PyObject_GetAttr(draw, n.offset_scale);
    bool ok = seeds && streams && item_numel && total && dtype && inner && scale && PyTuple_Check(seeds) && PyTuple_Check(streams) &&
              PyTuple_GET_SIZE(seeds) == PyTuple_GET_SIZE(streams) && PyTuple_GET_SIZE(seeds) >= 1 &&
              PyTuple_GET_SIZE(seeds) <= SKR_MAX_PHILOX_ITEMS;
    if (ok) {
        const Py_ssize_t count = PyTuple_GET_SIZE(seeds);
        for (Py_ssize_t i = 0; i < count; ++i) {
            table.seed[i] = PyLong_AsUnsignedLongLong(PyTuple_GET_ITEM(seeds, i));
            table.stream[i] = PyLong_AsUnsignedLongLong(PyTuple_GET_ITEM(streams, i));
        }
        table.n_items = (int32_t)count;
        table.item_numel = PyLong_AsLongLong(item_numel);
        *numel = PyLong_AsLongLong(total);
        table.dtype = THPDtype_Check(dtype) ? code_of(reinterpret_cast<THPDtype*>(dtype)->scalar_type) : 0;
        if (table.dtype < 0) table.dtype = 0;
        table.offset_inner = PyLong_AsLongLong(inner);
        table.offset_scale = (float)PyFloat_AsDouble(scale);
        table.reserved = 0;
        ok = !PyErr_Occurred();
    }
    PyErr_Clear();
    Py_XDECREF(seeds); Py_XDECREF(streams); Py_XDECREF(item_numel); Py_XDECREF(total); Py_XDECREF(dtype);
    Py_XDECREF(inner); Py_XDECREF(scale);
    return ok;
}

py::object hit(py::tuple roles, int n_inputs, int n_draws, py::object packed, py::object previous, py::dict plans, bool account) {
    const Py_ssize_t n = PyTuple_GET_SIZE(roles.ptr());
    if (n != n_inputs + n_draws || n_inputs < 1 || n_inputs > 32 || n_draws < 0 || n_draws > SKR_MAX_PHILOX) return py::none();
    py::list inputs(n_inputs);
    for (int i = 0; i < n_inputs; ++i) {
        PyObject* value = resolve(PyTuple_GET_ITEM(roles.ptr(), i), packed.ptr(), previous.ptr());
        if (!value) return py::none();
        PyList_SET_ITEM(inputs.ptr(), i, value);  // steals
    }
    static thread_local skr_philox tables[SKR_MAX_PHILOX];
    uintptr_t draws = 0;
    if (n_draws > 0) {
        PyObject* first = PyList_GET_ITEM(inputs.ptr(), 0);
        if (!THPVariable_Check(first)) return py::none();
        const int64_t expect = THPVariable_Unpack(first).numel();
        for (int d = 0; d < n_draws; ++d) {
            PyObject* draw = resolve(PyTuple_GET_ITEM(roles.ptr(), n_inputs + d), packed.ptr(), previous.ptr());
            int64_t numel = -1;
            const bool ok = draw && fill_draw(tables[d], draw, &numel) && numel == expect;
            Py_XDECREF(draw);
            if (!ok) return py::none();
        }
        draws = reinterpret_cast<uintptr_t>(tables);
    }
    return launch(plans, inputs, draws, account);
}

}  // namespace

PYBIND11_MODULE(_fast, m) {
    m.doc() = "skrample_b200: plan-cache hit path (torch tensors -> skr_plan_launch) in one call";
    m.def("bind", &bind, "hand over the address of skr_plan_launch");
    m.def("launch", &launch, py::arg("plans"), py::arg("inputs"), py::arg("draws") = 0, py::arg("account") = false);
    m.def("hit", &hit, "bind a step's tensors by role, fill the Philox key tables, launch");
}
