"""Process-global boundary-condition id registry (reference: boundary_condition_registry.py:6-30).

ids are handed out in order of BC *construction*, starting at 1; 0 = no boundary condition, 255 = solid / skip cell.
They end up in the uint8 ``bc_mask`` and must therefore match the reference's allocation exactly.
"""

FIRST_ID, LAST_ID = 1, 254  # uint8 with 0 (fluid) and 255 (solid) reserved


class BoundaryConditionRegistry:
    """`next_id`, `id_to_bc` and `bc_to_id` are public, as in the reference (scripts reset `next_id` between cases)."""

    def __init__(self):
        self.next_id = FIRST_ID
        self.id_to_bc, self.bc_to_id = {}, {}

    def register_boundary_condition(self, boundary_condition):
        assigned = self.next_id
        if assigned > LAST_ID:
            raise ValueError("more than 254 boundary conditions registered: ids must fit uint8 with 0 / 255 reserved")
        self.id_to_bc[assigned], self.bc_to_id[boundary_condition] = boundary_condition, assigned
        self.next_id = assigned + 1
        return assigned


boundary_condition_registry = BoundaryConditionRegistry()
