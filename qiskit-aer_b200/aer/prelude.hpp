// Force-included (-include) in front of the reference's UNMODIFIED pybind translation unit
// (/root/reference/qiskit_aer/backends/wrappers/bindings.cc) to build an Aer `controller_wrappers`
// module whose device="GPU" path runs on the B200 engine:
//
//   * AER_THRUST_SUPPORTED / AER_THRUST_GPU switch on the reference's own GPU branches
//     (src/controllers/aer_controller.hpp:290-330,636-647; src/simulators/circuit_executor.hpp:340-380);
//   * the reference's Thrust containers are kept out by pre-defining their include guards, and the class
//     names those branches instantiate are aliased: QubitVectorThrust<T> -> QubitVectorB200<T> and
//     DensityMatrixThrust<T> -> DensityMatrixB200<T> (ours); the unitary / superoperator "Thrust" names
//     fall back to the reference CPU classes (out of scope).
//
// This is the build-time equivalent of the two-line edit shown in INTEGRATION.md section 2; no reference
// file is modified or copied.
#ifndef _b200_aer_prelude_hpp_
#define _b200_aer_prelude_hpp_

#define AER_THRUST_SUPPORTED TRUE
#define AER_THRUST_GPU
#include <cuda_runtime.h>

#include "qubitvector_b200.hpp"
#include "densitymatrix_b200.hpp"

#define _qv_qubit_vector_thrust_hpp_
#define _qv_density_matrix_thrust_hpp_
#define _qv_unitary_matrix_thrust_hpp_
#define _qv_superoperator_thrust_hpp_
#include "framework/linalg/matrix_utils.hpp"
#include "simulators/density_matrix/densitymatrix.hpp"
#include "simulators/superoperator/superoperator.hpp"
#include "simulators/unitary/unitarymatrix.hpp"

namespace AER {
namespace QV {
template <typename T = double> using QubitVectorThrust = QubitVectorB200<T>;
template <typename T = double> using DensityMatrixThrust = DensityMatrixB200<T>;
template <typename T = double> using UnitaryMatrixThrust = UnitaryMatrix<T>;
template <typename T = double> using SuperoperatorThrust = Superoperator<T>;
}  // namespace QV
}  // namespace AER
#endif
