"""Set-up helpers of the Navier-Stokes stepper (field allocation, equilibrium initialisation, BC overlap check)."""

from xlb_b200._exports import export

export(globals(), __name__, {"nse_solver": ["create_nse_fields"], "initializers": ["initialize_eq"], "check_boundary_overlaps": ["check_bc_overlaps"]})
