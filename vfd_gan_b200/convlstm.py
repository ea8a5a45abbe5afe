"""ConvLSTM cell / unroll on the B200 kernels (reference: models/convlstm.py:6-169).

``ConvLSTMCell`` keeps the reference constructor, its single ``conv`` child (``nn.Conv2d`` holding
``conv.weight`` [4*hid, in+hid, kh, kw] with row blocks ordered i, f, o, g) and the
``forward(input_tensor, cur_state) -> (h_next, c_next)`` contract on fp32 NCHW tensors. The gate
convolution runs as a D=1 implicit GEMM on tcgen05 and the sigmoid/tanh cell update is one fused
kernel. ``init_hidden`` allocates on the module's device (the reference hard-codes ``.cuda()``,
models/convlstm.py:60-62).
"""
import torch
import torch.nn as nn

from . import ops


class _CellUpdateFn(torch.autograd.Function):
    """(gates fp32 channels-last [N,1,H,W,4*hid], c_cur fp32 [N,H,W,hid]) -> (h_next, c_next)."""

    @staticmethod
    def forward(ctx, gates, c_cur, hid):
        c_cur = c_cur.contiguous()
        h_next = torch.empty_like(c_cur)
        c_next = torch.empty_like(c_cur)
        act = torch.empty(*c_cur.shape[:-1], 4 * hid, dtype=torch.float32, device=c_cur.device)
        ops.convlstm_cell_fwd(gates, c_cur, h_next, c_next, act)
        ctx.save_for_backward(act, c_cur, c_next)
        ctx.gshape = gates.shape
        return h_next, c_next

    @staticmethod
    def backward(ctx, dh, dc):
        act, c_cur, c_next = ctx.saved_tensors
        dgates = torch.zeros(ctx.gshape, dtype=torch.bfloat16, device=act.device)
        dc_cur = torch.empty_like(c_cur)
        ops.convlstm_cell_bwd(act, c_cur, c_next, None if dh is None else dh.contiguous(),
                              None if dc is None else dc.contiguous(), dgates, dc_cur)
        return dgates, dc_cur, None


class ConvLSTMCell(nn.Module):
    def __init__(self, input_size, input_dim, hidden_dim, kernel_size, bias):
        super().__init__()
        self.height, self.width = input_size
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.kernel_size = kernel_size
        self.padding = kernel_size[0] // 2, kernel_size[1] // 2
        self.bias = bias
        self.conv = nn.Conv2d(in_channels=self.input_dim + self.hidden_dim, out_channels=4 * self.hidden_dim,
                              kernel_size=self.kernel_size, padding=self.padding, bias=self.bias)
        if any(k not in (1, 3) for k in kernel_size):
            raise NotImplementedError("ConvLSTMCell on B200 supports kernel extents 1 or 3")
        if hidden_dim % 2:
            raise NotImplementedError("ConvLSTMCell on B200 needs an even hidden_dim")

    def forward_cl(self, comb_cl, c_cur_cl):
        """comb_cl: channels-last bf16 [N,1,H,W,in+hid (padded)]; c_cur_cl fp32 [N,H,W,hid]."""
        gates = ops.ConvFn.apply(comb_cl, self.conv.weight, self.conv.bias, True, False)  # Conv2d == kd 1
        return _CellUpdateFn.apply(gates, c_cur_cl, self.hidden_dim)

    def forward(self, input_tensor, cur_state):
        h_cur, c_cur = cur_state
        combined = torch.cat([input_tensor, h_cur], dim=1)          # models/convlstm.py:46
        comb_cl = ops.PackFn.apply(combined.unsqueeze(2), 0)
        h_cl, c_cl = self.forward_cl(comb_cl, c_cur.permute(0, 2, 3, 1))
        return h_cl.permute(0, 3, 1, 2), c_cl.permute(0, 3, 1, 2)

    def init_hidden(self, batch_size):
        dev = self.conv.weight.device
        z = torch.zeros(batch_size, self.hidden_dim, self.height, self.width, device=dev)
        return (z, z.clone())


class ConvLSTM(nn.Module):
    """Layers x time unroll with zero initial state (models/convlstm.py:65-169)."""

    def __init__(self, input_size, input_dim, hidden_dim, kernel_size, num_layers, batch_first=False, bias=True,
                 return_all_layers=False):
        super().__init__()
        self._check_kernel_size_consistency(kernel_size)
        kernel_size = self._extend_for_multilayer(kernel_size, num_layers)
        hidden_dim = self._extend_for_multilayer(hidden_dim, num_layers)
        if not len(kernel_size) == len(hidden_dim) == num_layers:
            raise ValueError('Inconsistent list length.')
        self.height, self.width = input_size
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.kernel_size = kernel_size
        self.num_layers = num_layers
        self.batch_first = batch_first
        self.bias = bias
        self.return_all_layers = return_all_layers
        self.cell_list = nn.ModuleList([
            ConvLSTMCell(input_size=(self.height, self.width),
                         input_dim=self.input_dim if i == 0 else self.hidden_dim[i - 1],
                         hidden_dim=self.hidden_dim[i], kernel_size=self.kernel_size[i], bias=self.bias)
            for i in range(self.num_layers)])

    def forward(self, input_tensor, hidden_state=None):
        if not self.batch_first:
            input_tensor = input_tensor.permute(1, 0, 2, 3, 4)      # (t,b,c,h,w) -> (b,t,c,h,w)
        if hidden_state is not None:
            raise NotImplementedError()                              # as the reference, :120-121
        hidden_state = self._init_hidden(batch_size=input_tensor.size(0))
        layer_output_list, last_state_list = [], []
        cur_layer_input = input_tensor
        for layer_idx in range(self.num_layers):
            h, c = hidden_state[layer_idx]
            output_inner = []
            for t in range(cur_layer_input.size(1)):
                h, c = self.cell_list[layer_idx](input_tensor=cur_layer_input[:, t], cur_state=[h, c])
                output_inner.append(h)
            layer_output = torch.stack(output_inner, dim=1)
            cur_layer_input = layer_output
            layer_output_list.append(layer_output)
            last_state_list.append([h, c])
        if not self.return_all_layers:
            layer_output_list = layer_output_list[-1:]
            last_state_list = last_state_list[-1:]
        return layer_output_list, last_state_list

    def forward_cl(self, x_cl):
        """Channels-last unroll used inside the B200 nets: ``x_cl`` bf16 [N, T, H, W, C] (time on the depth
        axis) -> hidden states of the last layer, bf16 [N, T, H, W, hid]. Same recurrence as ``forward``
        (zero initial state, layers x time), without the NCHW round trips."""
        N, T, H, W, C = x_cl.shape
        if (H, W) != (self.height, self.width):
            raise RuntimeError(f"ConvLSTM built for {self.height}x{self.width} maps, got {H}x{W}")
        cur = x_cl
        for layer_idx, cell in enumerate(self.cell_list):
            cin = cell.input_dim
            if cin % 8 or cell.hidden_dim % 8 or cur.shape[-1] != cin:
                raise NotImplementedError("ConvLSTM.forward_cl needs input_dim and hidden_dim divisible by 8")
            h = torch.zeros(N, 1, H, W, cell.hidden_dim, dtype=torch.bfloat16, device=x_cl.device)
            c = torch.zeros(N, H, W, cell.hidden_dim, dtype=torch.float32, device=x_cl.device)
            outs = []
            fused = ops.lstm_step_fusable(cell.conv.weight, cin + cell.hidden_dim)
            if fused:   # gate conv + cell update in one kernel; the permuted gate operands are shared by the T steps
                w_perm, b_perm, kc = ops.lstm_gate_operands(cell.conv.weight, cell.conv.bias)
            for t in range(T):
                comb = torch.cat([cur[:, t:t + 1], h], dim=-1)        # models/convlstm.py:46
                if fused:
                    h, c = ops.LstmStepFn.apply(comb, c, cell.conv.weight, cell.conv.bias, w_perm, b_perm, kc)
                else:
                    h32, c = cell.forward_cl(comb, c)
                    h = h32.to(torch.bfloat16).unsqueeze(1)
                outs.append(h)
            cur = torch.cat(outs, dim=1)
        return cur

    def _init_hidden(self, batch_size):
        return [cell.init_hidden(batch_size) for cell in self.cell_list]

    @staticmethod
    def _check_kernel_size_consistency(kernel_size):
        if not (isinstance(kernel_size, tuple) or
                (isinstance(kernel_size, list) and all(isinstance(e, tuple) for e in kernel_size))):
            raise ValueError('`kernel_size` must be tuple or list of tuples')

    @staticmethod
    def _extend_for_multilayer(param, num_layers):
        return param if isinstance(param, list) else [param] * num_layers
