"""Host-side mirror of the reference interface (names follow PicoQuant.jl,
minus the Julia ``!``)."""
from .backends import (AbstractBackend, DSLBackend, Metrics, TensorStore, parse_dsl,
                       record_compute_costs, record_memory_costs)
from .circuit import (Circuit, gate_matrix, load_qasm_as_circuit,
                      load_qasm_as_circuit_from_file, transpile_circuit)
from .layer3 import (Edge, Node, TensorNetworkCircuit, add_gate, add_input, add_output,
                     convert_circuit_to_network, convert_qiskit_circ_to_network,
                     decompose_gate, edges, gate_data_from_matrix, gate_tensor, getedge,
                     getnode, inedges, inneighbours, neighbours, network_from_dict,
                     network_from_json, new_label, outedges, outneighbours, to_dict,
                     to_json, virtualedges, virtualneighbours)
from .layer2 import (calculate_mps_amplitudes, compress_bond, compress_tensor_chain,
                     contract_mps_tensor_network_circuit,
                     contract_tensor_network_circuit_with_compression, contract_network, contract_pair,
                     create_ncon_indices, decompose_tensor, full_wavefunction_contraction,
                     inorder_contraction, merge_common_bonds, random_contraction_plan,
                     sort_indices)
from .mps import MPSState
from .planner import (bgreedy, bgreedy_contraction, contraction_cost, greedy_plan, netcon,
                      netcon_contraction, plan_cost, sweep_plan)
from .slicing import (multi_index_partition, partition_network_on_virtual_bonds,
                      replace_with_view, slice_tensor_network)
from .algorithms import (create_ghz_preparation_circuit, create_qft_circuit, create_RQC,
                         create_simple_preparation_circuit, rqc_patterns)
