"""Two-class stand-in for the `qiskit_aer` Python package, only so that the reference's pybind layer
(src/framework/pybind_json.hpp:224-229, which imports these two names to recognise them in
`std::to_json(py::handle)`) can parse plain-dict noise models when qiskit itself is not installed."""
