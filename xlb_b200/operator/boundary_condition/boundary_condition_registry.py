"""Process-global boundary-condition id registry (reference: boundary_condition_registry.py:6-30).

ids are handed out in order of BC *construction*, starting at 1; 0 = no boundary condition, 255 = solid / skip cell.
They end up in the uint8 ``bc_mask`` and must therefore match the reference's allocation exactly.
"""


class BoundaryConditionRegistry:
    def __init__(self):
        self.id_to_bc = {}
        self.bc_to_id = {}
        self.next_id = 1  # 0 is reserved for no boundary condition

    def register_boundary_condition(self, boundary_condition):
        _id = self.next_id
        if _id > 254:
            raise ValueError("more than 254 boundary conditions registered: ids must fit uint8 with 0 / 255 reserved")
        self.next_id += 1
        self.id_to_bc[_id] = boundary_condition
        self.bc_to_id[boundary_condition] = _id
        return _id


boundary_condition_registry = BoundaryConditionRegistry()
