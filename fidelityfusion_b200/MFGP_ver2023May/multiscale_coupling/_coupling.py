"""Shared body of the gen-2023 fidelity couplings: high = rho * map(low) + res, with rho an fp32-pinned scalar
(reference Residual.py:14, matrix.py:67) and `map` the identity (Residual) or a chain of mode products
(Matrix_Mapping).  The variance transforms reuse the same maps, as in the reference."""
import torch


class RhoCoupling(torch.nn.Module):
    def _register_rho(self, value, trainable):
        self.rho = torch.nn.Parameter(torch.tensor(value, dtype=torch.float32), requires_grad=bool(trainable))

    def _map(self, low_fidelity):
        return low_fidelity

    def forward(self, low_fidelity, high_fidelity):
        """residual target of the high-fidelity GP"""
        return high_fidelity - self._map(low_fidelity) * self.rho

    def backward(self, low_fidelity, res):
        """high-fidelity prediction from the low-fidelity one and the predicted residual"""
        return self._map(low_fidelity) * self.rho + res

    def var_forward(self, low_fidelity_var, high_fidelity_var):
        return self.forward(low_fidelity_var, high_fidelity_var)

    def var_backward(self, low_fidelity_var, res_var):
        return self.backward(low_fidelity_var, res_var)
