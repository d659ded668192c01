"""Evaluate the serving function of a reference SavedModel straight from its `saved_model.pb` - TEST INFRASTRUCTURE.

The reference ships three trained exports (`src/main/resources/webroot/modeldata/neuralcf/{001,002}`, `MLPRec/005`).
TensorFlow cannot run here, but the graph TensorFlow serialised can be read: this module decodes the protobuf
without schemas (field numbers of saved_model.proto / meta_graph.proto / graph.proto / function.proto /
node_def.proto / attr_value.proto / tensor.proto), follows `serving_default` -> signature wrapper ->
`__inference__wrapped_model_*`, binds every resource argument of that function to its checkpoint tensor through the
export's own `__inference__traced_restore_*` function, and evaluates the function node by node in float32.

Everything is interpreted from the graph, not from the Python script: which placeholder and which table feed each
DenseFeatures block, the ~90 sparse ops of its embedding column (`to_sparse_input` with ignore value -1,
SparseReshape, SparseFillEmptyRows, Unique, ResourceGather, SparseSegmentMean, Select) and its range asserts, the
order of the concatenation, which kernel / bias every MatMul / BiasAdd reads, the activations, the Dot / Squeeze
tail, the output tensor.  What is restated is only the meaning of each TensorFlow op in numpy (`node_value`).
`run(..., full=False)` replaces a single-column DenseFeatures block by "row id of the table its ResourceGather
reads"; both give the same 3 x 514 outputs bit for bit.

Used by `tests/golden/make_savedmodel_graph_vectors.py` (writes `tests/golden/savedmodel_graph_vectors.json`) and
`tests/test_oracle_golden.py`; never imported by the product.
"""
import os
import struct

import numpy as np


# ---- schema-less protobuf ---------------------------------------------------------------------------------
def _varint(b, p):
    r = s = 0
    while True:
        c = b[p]
        p += 1
        r |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return r, p


def fields(b):
    p, out = 0, []
    while p < len(b):
        k, p = _varint(b, p)
        f, w = k >> 3, k & 7
        if w == 0:
            v, p = _varint(b, p)
        elif w == 2:
            n, p = _varint(b, p)
            v = b[p:p + n]
            p += n
        elif w == 5:
            v = b[p:p + 4]
            p += 4
        elif w == 1:
            v = b[p:p + 8]
            p += 8
        else:
            raise ValueError("wire type %d" % w)
        out.append((f, w, v))
    return out


def get(fs, f):
    return [v for (ff, _, v) in fs if ff == f]


def _packed_varints(v):
    if isinstance(v, int):
        return [v]
    out, p = [], 0
    while p < len(v):
        x, p = _varint(v, p)
        out.append(x)
    return out


def _signed(x, bits=64):
    return x - (1 << bits) if x >= 1 << (bits - 1) else x


def tensor_proto(buf):
    """TensorProto -> numpy (the dtypes these graphs use)."""
    f = fields(buf)
    dtype = get(f, 1)[0]
    shape = [_signed(get(fields(d), 1)[0]) if get(fields(d), 1) else 0 for d in get(fields(get(f, 2)[0]), 2)] if get(f, 2) else []
    content = get(f, 4)
    if dtype == 7:                                   # DT_STRING
        return np.array([s.decode() for s in get(f, 8)], dtype=object).reshape(shape)
    np_t = {1: np.float32, 3: np.int32, 9: np.int64, 10: np.bool_}[dtype]
    if content:
        return np.frombuffer(content[0], dtype=np_t).reshape(shape).copy()
    if dtype == 1:
        vals = [x for v in get(f, 5) for x in (struct.unpack("<%df" % (len(v) // 4), v) if isinstance(v, bytes) else [v])]
    elif dtype == 3:
        vals = [_signed(x) for v in get(f, 7) for x in _packed_varints(v)]
    elif dtype == 9:
        vals = [_signed(x) for v in get(f, 10) for x in _packed_varints(v)]
    else:
        vals = [x for v in get(f, 11) for x in _packed_varints(v)]
    n = int(np.prod(shape)) if shape else 1
    if len(vals) == 1 and n > 1:
        vals = vals * n
    return np.array(vals, dtype=np_t).reshape(shape)


class Node:
    def __init__(self, buf):
        f = fields(buf)
        self.name = get(f, 1)[0].decode()
        self.op = get(f, 2)[0].decode()
        self.inputs = [i.decode() for i in get(f, 3)]
        self.attr = {}
        for a in get(f, 5):
            e = fields(a)
            self.attr[get(e, 1)[0].decode()] = fields(get(e, 2)[0]) if get(e, 2) else []

    def data_inputs(self):
        return [i for i in self.inputs if not i.startswith("^")]

    def func(self, key="f"):
        return get(fields(get(self.attr[key], 10)[0]), 1)[0].decode()

    def b(self, key, default=False):
        v = get(self.attr.get(key, []), 5)
        return bool(v[0]) if v else default

    def i(self, key, default=0):
        v = get(self.attr.get(key, []), 3)
        return _signed(v[0]) if v else default

    def dtype(self, key):
        return {1: np.float32, 3: np.int32, 9: np.int64, 10: np.bool_}[get(self.attr[key], 6)[0]]

    def ints(self, key):
        lst = get(self.attr.get(key, []), 1)
        return [_signed(x) for v in get(fields(lst[0]), 3) for x in _packed_varints(v)] if lst else []


class Function:
    def __init__(self, buf):
        f = fields(buf)
        sig = fields(get(f, 1)[0])
        self.name = get(sig, 1)[0].decode()
        self.args = [get(fields(a), 1)[0].decode() for a in get(sig, 2)]
        self.outs = [get(fields(a), 1)[0].decode() for a in get(sig, 3)]
        self.nodes = {}
        for n in get(f, 3):
            nd = Node(n)
            self.nodes[nd.name] = nd
        self.ret = {get(fields(r), 1)[0].decode(): get(fields(r), 2)[0].decode() for r in get(f, 4)}


class ServingGraph:
    """The `serving_default` computation of one export, with its variables bound."""

    def __init__(self, savedmodel_dir, read_variables):
        sm = fields(open(os.path.join(savedmodel_dir, "saved_model.pb"), "rb").read())
        mg = fields(get(sm, 2)[0])
        gd = fields(get(mg, 2)[0])
        top = [Node(n) for n in get(gd, 1)]
        self.funcs = {}
        for fb in get(fields(get(gd, 2)[0]), 1):
            fn = Function(fb)
            self.funcs[fn.name] = fn
        calls = [n for n in top if n.op in ("StatefulPartitionedCall", "PartitionedCall")]
        sig_call = [n for n in calls if n.func().startswith("__inference_signature_wrapper")]
        assert len(sig_call) == 1
        wrapper = self.funcs[sig_call[0].func()]
        inner = [n for n in wrapper.nodes.values() if n.op == "StatefulPartitionedCall"]
        assert len(inner) == 1 and inner[0].data_inputs() == wrapper.args, "the wrapper forwards its arguments in order"
        self.fn = self.funcs[inner[0].func()]
        assert self.fn.name.startswith("__inference__wrapped_model")
        bound = dict(zip(self.fn.args, sig_call[0].data_inputs()))      # function argument -> top-level node name
        self.placeholders = {a: v[len("serving_default_"):] for a, v in bound.items() if v.startswith("serving_default_")}
        var_of_arg = {a: v for a, v in bound.items() if a not in self.placeholders}
        # variable node -> checkpoint key, read off the export's own restore function
        rest_call = [n for n in calls if n.func().startswith("__inference__traced_restore")][0]
        rest = self.funcs[rest_call.func()]
        node_of_arg = dict(zip(rest.args, rest_call.data_inputs()))
        names = tensor_proto(get(rest.nodes["RestoreV2/tensor_names"].attr["value"], 8)[0]).reshape(-1)
        key_of_var = {}
        for n in rest.nodes.values():
            if n.op != "AssignVariableOp":
                continue
            res, val = n.data_inputs()
            src = rest.nodes[val.split(":")[0]].data_inputs()[0]         # Identity_k <- RestoreV2:tensors:k
            if not src.startswith("RestoreV2:tensors:"):
                continue
            key_of_var[node_of_arg[res]] = names[int(src.rsplit(":", 1)[1])]
        ckpt = read_variables(os.path.join(savedmodel_dir, "variables"))
        # checkpoint keys are object-graph paths: "<path>/.ATTRIBUTES/VARIABLE_VALUE", "/" inside a name escaped as ".S"
        strip = lambda k: (k[:-len("/.ATTRIBUTES/VARIABLE_VALUE")] if k.endswith("/.ATTRIBUTES/VARIABLE_VALUE") else k).replace(".S", "/")
        self.other_resources = {a: v for a, v in var_of_arg.items() if v not in key_of_var}   # e.g. vocabulary hash tables
        var_of_arg = {a: v for a, v in var_of_arg.items() if v in key_of_var}
        self.variables = {a: np.asarray(ckpt[strip(key_of_var[v])], dtype=np.float32) for a, v in var_of_arg.items()}
        self.variable_names = {a: (v, strip(key_of_var[v])) for a, v in var_of_arg.items()}
        self.top_nodes = {n.name: n for n in top}
        self.trace = []                                                  # (node, op) in evaluation order

    def vocabulary_tables(self):
        """{column (lower case, from the argument name): (keys, values, default)} for every vocabulary-list column:
        the keys / values the export's initialisers import into its hash tables and the lookup's default value."""
        top = self.top_nodes
        const = lambda name: tensor_proto(get(top[name].attr["value"], 8)[0])
        imported = {}
        for n in top.values():
            if n.op in ("StatefulPartitionedCall", "PartitionedCall") and "f" in n.attr:
                fn = self.funcs[n.func()]
                if any(m.op == "LookupTableImportV2" for m in fn.nodes.values()):
                    t, k, v = n.data_inputs()
                    imported[t] = (const(k), const(v))
        out = {}
        for arg, node in self.other_resources.items():
            col = arg.split("dense_features_")[1].split("_indicator")[0].split("_embedding")[0]
            if arg.endswith("table_handle"):
                out.setdefault(col, {})["keys"], out[col]["values"] = imported[node]
            elif arg.endswith("default_value"):
                out.setdefault(col, {})["default"] = const(node)
        return out

    def dense_features_order(self, scope_suffix="dense_features/concat"):
        """Column names in the order a multi-column DenseFeatures block concatenates them."""
        n = [m for m in self.fn.nodes.values() if m.op == "ConcatV2" and m.name.endswith(scope_suffix)]
        assert len(n) == 1
        return [i.split(":")[0].split("/")[-2] for i in n[0].data_inputs()[:-1]]

    def dense_features_blocks(self):
        """{scope: (placeholder, variable node name, number of nodes)} for every DenseFeatures block of the function."""
        out = {}
        for n in self.fn.nodes.values():
            if n.op != "ResourceGather":
                continue
            scope = "/".join(n.name.split("/")[:2])
            inside = [m for m in self.fn.nodes.values() if m.name.startswith(scope + "/")]
            fed = sorted({self.placeholders[i] for m in inside for i in m.data_inputs() if i in self.placeholders})
            assert len(fed) == 1 and scope not in out
            out[scope] = (fed[0], self.variable_names[n.data_inputs()[0]][0], len(inside))
        return out

    # ---- evaluation -----------------------------------------------------------------------------------------
    def run(self, feeds, full=True):
        """feeds: {placeholder name (e.g. "movieId"): 1-D array}.  Returns the function's output array.
        full=True executes every node of the function, the ~90 sparse ops of each DenseFeatures block included
        (`to_sparse_input` -> SparseReshape -> SparseFillEmptyRows -> Unique -> ResourceGather -> SparseSegmentMean
        -> Select) and its range asserts; full=False takes a single-column block as "row id of the table its
        ResourceGather reads" (the shortcut the first version of the golden vectors used - kept as a cross-check)."""
        memo = {}
        self.trace = []

        def shortcut(scope):
            inside = [n for n in self.fn.nodes.values() if n.name.startswith(scope + "/")]
            gathers = [n for n in inside if n.op == "ResourceGather"]
            assert len(gathers) == 1, "%s: expected one embedding column, found %d" % (scope, len(gathers))
            table = self.variables[gathers[0].data_inputs()[0]]
            fed = {i for n in inside for i in n.data_inputs() if i in self.placeholders}
            assert len(fed) == 1, "%s reads placeholders %s" % (scope, sorted(fed))
            ids = np.asarray(feeds[self.placeholders[fed.pop()]]).astype(np.int64).reshape(-1)
            return table[ids]

        def node_value(n):
            x = [ev(i) for i in n.data_inputs()]
            op = n.op
            if op == "Const":
                return tensor_proto(get(n.attr["value"], 8)[0])
            if op == "ReadVariableOp":
                return self.variables[x[0]]
            if op in ("Identity", "StopGradient"):
                return x[0]
            if op == "ConcatV2":
                return np.concatenate(x[:-1], axis=int(x[-1]))
            if op == "MatMul":
                a = x[0].T if n.b("transpose_a") else x[0]
                b = x[1].T if n.b("transpose_b") else x[1]
                return (a.astype(np.float32) @ b.astype(np.float32)).astype(np.float32)
            if op == "BiasAdd":
                return (x[0] + x[1]).astype(np.float32)
            if op == "Relu":
                return np.maximum(x[0], np.float32(0))
            if op == "Sigmoid":
                return (1.0 / (1.0 + np.exp(-x[0].astype(np.float64)))).astype(np.float32)
            if op == "ExpandDims":
                return np.expand_dims(x[0], int(x[1]))
            if op == "BatchMatMulV2":
                a = np.swapaxes(x[0], -1, -2) if n.b("adj_x") else x[0]
                b = np.swapaxes(x[1], -1, -2) if n.b("adj_y") else x[1]
                return np.matmul(a.astype(np.float32), b.astype(np.float32)).astype(np.float32)
            if op == "Squeeze":
                dims = n.ints("squeeze_dims")
                return np.squeeze(x[0], axis=tuple(dims)) if dims else np.squeeze(x[0])
            # ---- the sparse embedding lookup of a DenseFeatures block ------------------------------------
            if op in ("NotEqual", "Less", "GreaterEqual"):
                return {"NotEqual": np.not_equal, "Less": np.less, "GreaterEqual": np.greater_equal}[op](x[0], x[1])
            if op == "Where":
                return np.argwhere(x[0]).astype(np.int64)
            if op == "GatherNd":
                return x[0][tuple(np.asarray(x[1]).T)]
            if op == "Shape":
                return np.array(np.shape(x[0]), dtype=n.dtype("out_type") if "out_type" in n.attr else np.int32)
            if op == "Cast":
                return np.asarray(x[0]).astype(n.dtype("DstT"))
            if op == "All":
                return np.all(x[0])
            if op == "If":                                   # the Assert guards of the identity column
                if not bool(x[0]):
                    raise ValueError("assertion of the graph failed: %s" % n.name)
                return x[1]
            if op == "Slice":
                begin, size = [int(v) for v in x[1]], [int(v) for v in x[2]]
                return x[0][tuple(slice(b, None if z < 0 else b + z) for b, z in zip(begin, size))]
            if op == "Prod":
                return np.prod(x[0], axis=tuple(np.atleast_1d(x[1]).tolist())).astype(x[0].dtype)
            if op == "GatherV2":
                return np.take(x[0], x[1], axis=int(x[2]))
            if op == "Pack":
                return np.stack([np.asarray(v) for v in x], axis=n.i("axis", 0))
            if op == "Reshape":
                return np.reshape(x[0], [int(v) for v in np.atleast_1d(x[1])])
            if op == "SparseReshape":
                idx, shape, new = x[0], [int(v) for v in x[1]], [int(v) for v in x[2]]
                total = int(np.prod(shape))
                if -1 in new:
                    new[new.index(-1)] = total // int(-np.prod(new))
                lin = np.ravel_multi_index(tuple(idx.T), shape) if len(idx) else np.zeros(0, np.int64)
                out = np.stack(np.unravel_index(lin, new), axis=1).astype(np.int64) if len(idx) else np.zeros((0, len(new)), np.int64)
                return {"output_indices": out, "output_shape": np.array(new, np.int64)}
            if op == "SparseFillEmptyRows":
                idx, vals, shape, default = x[0], x[1], [int(v) for v in x[2]], x[3]
                rows = shape[0]
                present = np.zeros(rows, bool)
                present[idx[:, 0]] = True
                add = np.flatnonzero(~present)
                all_idx = np.concatenate([idx, np.stack([add, np.zeros_like(add)], axis=1)], axis=0)
                all_val = np.concatenate([vals, np.full(len(add), default, dtype=vals.dtype)])
                order = np.lexsort((all_idx[:, 1], all_idx[:, 0]))
                return {"output_indices": all_idx[order].astype(np.int64), "output_values": all_val[order],
                        "empty_row_indicator": ~present, "reverse_index_map": np.argsort(order)[:len(idx)].astype(np.int64)}
            if op == "StridedSlice":
                begin, end, stride = ([int(v) for v in np.atleast_1d(a)] for a in x[1:4])
                bm, em, sm = n.i("begin_mask"), n.i("end_mask"), n.i("shrink_axis_mask")
                assert n.i("ellipsis_mask") == 0 and n.i("new_axis_mask") == 0
                key = []
                for d in range(len(begin)):
                    if sm >> d & 1:
                        key.append(begin[d])
                    else:
                        key.append(slice(None if bm >> d & 1 else begin[d], None if em >> d & 1 else end[d], stride[d]))
                return np.asarray(x[0])[tuple(key)]
            if op == "Unique":
                y, first, inv = np.unique(x[0], return_index=True, return_inverse=True)
                order = np.argsort(first)                    # TF keeps first-occurrence order
                rank = np.empty_like(order)
                rank[order] = np.arange(len(order))
                return {"y": y[order], "idx": rank[inv].astype(np.int32)}
            if op == "ResourceGather":
                return self.variables[x[0]][np.asarray(x[1]).astype(np.int64)]
            if op == "SparseSegmentMean":
                data, indices, seg = x[0], np.asarray(x[1]).astype(np.int64), np.asarray(x[2]).astype(np.int64)
                out = np.zeros((int(seg.max()) + 1 if len(seg) else 0,) + data.shape[1:], np.float32)
                cnt = np.zeros(out.shape[0], np.float32)
                np.add.at(out, seg, data[indices])
                np.add.at(cnt, seg, 1)
                return (out / np.maximum(cnt, 1)[:, None]).astype(np.float32)
            if op == "Tile":
                return np.tile(x[0], [int(v) for v in x[1]])
            if op == "ZerosLike":
                return np.zeros_like(x[0])
            if op in ("Select", "SelectV2"):
                return np.where(x[0], x[1], x[2])
            raise NotImplementedError("op %s (%s)" % (op, n.name))

        def ev(ref):
            parts = ref.split(":")
            name = parts[0]
            if name in self.placeholders:
                return np.asarray(feeds[self.placeholders[name]])
            if name in self.variables:
                return name                                              # a resource handle
            if name not in memo:
                n = self.fn.nodes[name]
                for c in n.inputs:                                       # control dependencies first: the range asserts
                    if c.startswith("^") and c[1:] in self.fn.nodes and c[1:] not in memo:
                        ev(c[1:])
                if not full and name.endswith("/concat/concat") and name.split("/")[-3].startswith("dense_features"):
                    memo[name] = shortcut(name[:-len("/concat/concat")])
                else:
                    memo[name] = node_value(n)
                self.trace.append((name, n.op))
            v = memo[name]
            return v[parts[1]] if isinstance(v, dict) else v

        assert len(self.fn.outs) == 1
        return ev(self.fn.ret[self.fn.outs[0]])
