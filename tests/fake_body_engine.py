"""TEST AID: an `engine.rewrite_bodies` look-alike backed by the g++ host build of the device machine,
so the host-side logic of responses.py can be exercised on the CPU suite."""
import host_machine as hm


class HostBodyEngine:
    def __init__(self, plans, fast=False):
        self.packed = plans.packed()
        self.fast = fast

    def rewrite_bodies(self, bodies, plan_idx, slot_cap=None, with_matched=False):
        plans, ops, blob = self.packed
        rows = []
        for raw, pi in zip(bodies, plan_idx):
            st = hm.FAST_IRREGULAR
            if self.fast:
                st, out, need = hm.rewrite_body_fast(raw, plans, ops, blob, int(pi))
            if st == hm.FAST_IRREGULAR:
                st, out, need = hm.rewrite_body(raw, plans, ops, blob, int(pi))
            out = out if st == 0 else b""
            rows.append((st, out, hm.last_matched(), hm.last_root_kind()) if with_matched else (st, out))
        return rows

    def documents_error_detail(self, docs, text_stride: int = 4096):
        return [hm.error_detail(bytes(d), text_stride) for d in docs]
