"""unitair_b200 -- B200-native (sm_100a) state-vector engine behind unitair's gate-application
API.  Drop-in for the hot path of qcware/qcware-unitair:

    import unitair_b200 as unitair
    psi = unitair.simulation.apply_operator(operator=h, qubits=(0,), state=psi)

States are plain torch complex CUDA tensors of size (*batch_dims, 2**n); see
simulation/operations.py and states/innerprod.py for the mirrored functions and
circuit.py / sharded.py for the circuit-level and multi-GPU extensions.
"""
from . import states
from . import simulation
from . import gates
from . import initializations
from . import circuit
from . import hoststream
from . import batch
from .batch import shard_batch, all_reduce_gradients
from .hoststream import HostCircuitStream, ShardedHostStream

from .states.shapes import StateLayout
from .initializations import unit_vector, uniform_superposition, rand_state

from .states import count_qubits, hilbert_space_dim
from .states import diag_expectation_value
from .states import inner_product, norm_squared, abs_squared

VECTOR_LAYOUT = states.shapes.StateLayout.VECTOR
TENSOR_LAYOUT = states.shapes.StateLayout.TENSOR

__version__ = "0.1.0"
