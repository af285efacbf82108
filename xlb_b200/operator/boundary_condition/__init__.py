"""Boundary conditions (namespace of reference xlb/operator/boundary_condition; GradsApproximationBC is out of scope)."""

from xlb_b200._exports import export

export(
    globals(),
    __name__,
    {
        "helper_functions_bc": ["HelperFunctionsBC"],
        "boundary_condition": ["BoundaryCondition", "ImplementationStep"],
        "boundary_condition_registry": ["BoundaryConditionRegistry"],
        "bc_equilibrium": ["EquilibriumBC"],
        "bc_do_nothing": ["DoNothingBC"],
        "bc_halfway_bounce_back": ["HalfwayBounceBackBC"],
        "bc_fullway_bounce_back": ["FullwayBounceBackBC"],
        "bc_zouhe": ["ZouHeBC"],
        "bc_regularized": ["RegularizedBC"],
        "bc_extrapolation_outflow": ["ExtrapolationOutflowBC"],
    },
)
