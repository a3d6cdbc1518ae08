from .operations import apply_phase
from .operations import apply_operator
from .operations import apply_to_qubits
from .operations import apply_all_qubits
from .operations import swap, roll_qubits, permute_qubits
from .operations import act_first_qubits
