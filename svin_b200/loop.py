"""Host-side mirror of the loop-closure query calls of pose_graph over the C ABI (svin_loop_*, include/svin_b200.h):
DBoW2 `voc->transform`, `db.add`, `db.query` (pose_graph/src/pose_graph/PoseGraph.cpp:170-224) and
`Keyframe::searchByBRIEFDes` (pose_graph/src/pose_graph/Keyframe.cpp:288-306).  No CPU path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class LoopEngine:
    def __init__(self, first_child, num_children, descriptor, weight, word_id, device=0, rank=0, world=1):
        self._lib = capi.load()
        c = np.ascontiguousarray
        self._keep = (c(first_child, np.int32), c(num_children, np.int32), c(descriptor, np.uint8).reshape(-1, 32),
                      c(weight, np.float64), c(word_id, np.int32))
        v = capi.SvinVocabulary()
        v.num_nodes = len(self._keep[0])
        v.first_child = self._keep[0].ctypes.data_as(capi.c_int32_p)
        v.num_children = self._keep[1].ctypes.data_as(capi.c_int32_p)
        v.descriptor = self._keep[2].ctypes.data_as(capi.c_uint8_p)
        v.weight = self._keep[3].ctypes.data_as(capi.c_double_p)
        v.word_id = self._keep[4].ctypes.data_as(capi.c_int32_p)
        self._ctx = C.c_void_p()
        capi.check(self._lib.svin_loop_create(device, C.byref(v), rank, world, C.byref(self._ctx)), self._lib)

    def close(self):
        if self._ctx:
            self._lib.svin_loop_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _pack(images):
        counts = np.array([len(x) for x in images], dtype=np.int32)
        desc = (np.ascontiguousarray(np.concatenate([np.asarray(x, np.uint8).reshape(-1, 32) for x in images]))
                if len(images) and counts.sum() else np.zeros((0, 32), np.uint8))
        return desc, counts

    def transform(self, images):
        """-> [(word ids ascending, L1-normalised TF-IDF values)] per image."""
        desc, counts = self._pack(images)
        n = int(counts.sum())
        ids, vals, nw = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1)), np.zeros(max(len(images), 1), np.int32)
        capi.check(self._lib.svin_loop_transform(self._ctx, len(images), desc.ctypes.data_as(capi.c_uint8_p),
                                                 counts.ctypes.data_as(capi.c_int32_p), ids.ctypes.data_as(capi.c_int32_p),
                                                 vals.ctypes.data_as(capi.c_double_p), nw.ctypes.data_as(capi.c_int32_p)),
                   self._lib)
        off = np.concatenate([[0], np.cumsum(counts)])
        return [(ids[off[i]:off[i] + nw[i]].copy(), vals[off[i]:off[i] + nw[i]].copy()) for i in range(len(images))]

    def add(self, images):
        desc, counts = self._pack(images)
        capi.check(self._lib.svin_loop_add(self._ctx, len(images), desc.ctypes.data_as(capi.c_uint8_p),
                                           counts.ctypes.data_as(capi.c_int32_p)), self._lib)

    def query(self, features, max_results=4, max_id=-1):
        """db.query(bowVec, ret, max_results, max_id) -> [(entry id, score)] best first."""
        f = np.ascontiguousarray(np.asarray(features, np.uint8).reshape(-1, 32))
        ids, sc, n = np.zeros(64, np.int32), np.zeros(64), C.c_int32()
        capi.check(self._lib.svin_loop_query(self._ctx, f.ctypes.data_as(capi.c_uint8_p), len(f), max_results, max_id,
                                             ids.ctypes.data_as(capi.c_int32_p), sc.ctypes.data_as(capi.c_double_p),
                                             C.byref(n)), self._lib)
        return [(int(ids[k]), float(sc[k])) for k in range(n.value)]

    def brief_search(self, window_desc, old_desc):
        w = np.ascontiguousarray(np.asarray(window_desc, np.uint8).reshape(-1, 32))
        o = np.ascontiguousarray(np.asarray(old_desc, np.uint8).reshape(-1, 32))
        idx, dist, st = np.zeros(max(len(w), 1), np.int32), np.zeros(max(len(w), 1), np.int32), np.zeros(max(len(w), 1), np.uint8)
        capi.check(self._lib.svin_loop_brief_search(self._ctx, w.ctypes.data_as(capi.c_uint8_p), len(w),
                                                    o.ctypes.data_as(capi.c_uint8_p), len(o),
                                                    idx.ctypes.data_as(capi.c_int32_p), dist.ctypes.data_as(capi.c_int32_p),
                                                    st.ctypes.data_as(capi.c_uint8_p)), self._lib)
        return idx[:len(w)], dist[:len(w)], st[:len(w)]

    def stats(self):
        a, b, ms = C.c_int32(), C.c_int32(), C.c_double()
        capi.check(self._lib.svin_loop_stats(self._ctx, C.byref(a), C.byref(b), C.byref(ms)), self._lib)
        return dict(entries_total=a.value, entries_local=b.value, last_device_ms=ms.value)


def merge_shards(results_per_rank, max_results=4):
    """The exchange step of the sharded query: every rank's best (entry, score) pairs -> global best, the reference's order
    (descending score, ties by entry id)."""
    allr = [r for rr in results_per_rank for r in rr]
    allr.sort(key=lambda t: (-t[1], t[0]))
    return allr[:max_results]
