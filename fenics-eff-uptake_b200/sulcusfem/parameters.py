"""Host-side parameters of the sulcus transport model (mirror of the reference's parameters.py API).

Same names, argument meaning and error behaviour as the reference module so that study drivers can
switch their import: ``Parameters`` (``parameters.py:92-334``: dimensional inputs, ``validate()``,
``nondim()``, ``get_mesh_generator_params()``), ``StepUptakeOpen`` (``parameters.py:24-84``) and the
geometry catalogue ``create_geometry_variations`` (``parameters.py:342-447``).  Pure host
configuration: nothing here runs on the hot path, except that ``StepUptakeOpen`` additionally offers
a vectorised ``mu_at`` so the Robin coefficient can be tabulated at the boundary nodes without a
per-node Python callback.
"""
from __future__ import annotations

import warnings

import numpy as np

from .fem import UserExpression

_MODES = ('adv-diff', 'no-adv', 'no-uptake')


class StepUptakeOpen(UserExpression):
    """Step Robin coefficient mu(x) on y=0: ``mu_base`` outside the mouth ``[xL, xR]``, blended to
    ``mu_eff_target`` inside with a logistic ramp of width ``L_c`` next to each mouth edge."""

    def __init__(self, mu_base, mu_eff_target, sulcus_left_x, sulcus_right_x, L_c=None, Gamma=5.0, **kwargs):
        super().__init__(**kwargs)
        self.xL, self.xR = float(sulcus_left_x), float(sulcus_right_x)
        self.w = float(self.xR - self.xL)
        if self.w <= 0:
            raise ValueError(f"sulcus_right_x must be > sulcus_left_x (got w={self.w})")
        self.mu_base, self.mu_open, self.Gamma = float(mu_base), float(mu_eff_target), float(Gamma)
        ramp = 0.1 * self.w if L_c is None else float(L_c)
        self.L_c = max(0.0, min(ramp, 0.49 * self.w))

    def mu_at(self, x):
        """Vectorised mu(x) (same branches as the per-point ``eval``)."""
        x = np.asarray(x, dtype=np.float64)
        inside = (x >= self.xL) & (x <= self.xR)
        alpha = np.ones_like(x)
        if self.L_c > 0.0:
            d = np.minimum(x - self.xL, self.xR - x)
            ramp = inside & (d < self.L_c)
            z = np.where(ramp, d / self.L_c, 0.0)
            alpha = np.where(ramp, 1.0 / (1.0 + np.exp(-self.Gamma * (z - 0.5))), 1.0)
        return np.where(inside, (1.0 - alpha) * self.mu_base + alpha * self.mu_open, self.mu_base)

    def _alpha_at(self, x):
        if x < self.xL or x > self.xR:
            return 0.0
        return float((self.mu_at(np.array([x]))[0] - self.mu_base) / (self.mu_open - self.mu_base)) \
            if self.mu_open != self.mu_base else 1.0

    def eval(self, values, x):
        values[0] = float(self.mu_at(np.array([x[0]]))[0])

    def value_shape(self):
        return ()


class Parameters:
    """User parameters; call ``validate()`` then ``nondim()`` (lengths scaled by the channel height)."""

    MU_DIM_ADV_DIFF = 0.0003
    MU_DIM_NO_ADV = 0.0003
    MU_DIM_NO_UPTAKE = 0
    VALID_MODES = set(_MODES)
    VISCOSITY = 1.0
    RHO = 1.0

    def __init__(self, mode='adv-diff', L_dim=10.0, H_dim=1.0, sulci_n=1, sulci_w_dim=0.5, sulci_h_dim=1.0,
                 mesh_size_dim=0.02, refinement_factor=1, U_ref_dim=0.012, D_dim=0.0003):
        if mode not in self.VALID_MODES:
            raise ValueError(f"Mode must be one of {self.VALID_MODES}, got '{mode}'")
        self.mode = mode
        self.L_dim, self.H_dim = L_dim, H_dim
        self.sulci_n, self.sulci_w_dim, self.sulci_h_dim = sulci_n, sulci_w_dim, sulci_h_dim
        self.mesh_size_dim, self.refinement_factor = mesh_size_dim, refinement_factor
        self.U_ref_dim, self.D_dim = U_ref_dim, D_dim
        self.mu_dim = {'adv-diff': self.MU_DIM_ADV_DIFF, 'no-adv': self.MU_DIM_NO_ADV,
                       'no-uptake': self.MU_DIM_NO_UPTAKE}[mode]

    # ------------------------------------------------------------------ validation
    @staticmethod
    def _validate_positive(value, name):
        if value <= 0:
            raise ValueError(f"{name} must be > 0, got {value}")

    @staticmethod
    def _validate_non_negative(value, name):
        if value < 0:
            raise ValueError(f"{name} cannot be negative, got {value}")

    def validate(self):
        self._validate_positive(self.L_dim, 'Domain length')
        self._validate_positive(self.H_dim, 'Domain height')
        for v, nm in ((self.sulci_n, 'Number of sulci'), (self.sulci_h_dim, 'Sulcus height'), (self.sulci_w_dim, 'Sulci width')):
            self._validate_non_negative(v, nm)
        if self.sulci_n > 0:
            self._validate_positive(self.sulci_h_dim, 'Sulcus height (when sulci defined)')
            self._validate_positive(self.sulci_w_dim, 'Sulcus width (when sulci defined)')
            if self.sulci_w_dim * self.sulci_n >= self.L_dim:
                raise ValueError("Total sulcus width must be less than domain length.")
        self._validate_positive(self.mesh_size_dim, 'Mesh size')
        if not isinstance(self.refinement_factor, int) or self.refinement_factor < 1:
            raise ValueError("Refinement factor must be an integer ≥ 1.")
        smallest = min(self.L_dim, self.H_dim)
        if self.mesh_size_dim > smallest / 10:
            warnings.warn(f"Mesh size ({self.mesh_size_dim}) is large relative to domain.")
        if self.mesh_size_dim < smallest / 1000:
            warnings.warn(f"Mesh size ({self.mesh_size_dim}) is very small - may be slow.")
        if self.mode in ('adv-diff', 'no-uptake'):
            self._validate_non_negative(self.U_ref_dim, 'Reference velocity')
        self._validate_non_negative(self.D_dim, 'Diffusion coefficient')
        if self.mode == 'no-adv' and self.D_dim <= 0:
            raise ValueError("Diffusion coefficient must be > 0 for diffusion-only mode.")
        if self.mode == 'no-uptake':
            if self.mu_dim != 0:
                warnings.warn("Setting mu to 0 for no-uptake mode.")
                self.mu_dim = 0
        else:
            self._validate_non_negative(self.mu_dim, 'Uptake parameter')

    # ------------------------------------------------------------------ scaling
    def nondim(self):
        ref = self.L_ref = self.H_dim
        self.L, self.H = self.L_dim / ref, self.H_dim / ref
        self.sulci_h, self.sulci_w = self.sulci_h_dim / ref, self.sulci_w_dim / ref
        self.mesh_size = self.mesh_size_dim / ref
        if self.mode == 'no-adv':
            self.D, self.U_ref, self.Pe, self.Re = 1.0, 0.0, None, None
        else:
            self.Pe = (self.U_ref_dim * self.H_dim) / self.D_dim
            self.D = 1.0 / self.Pe
            self.Re = (self.RHO * self.U_ref_dim * self.L_ref) / self.VISCOSITY
            self.U_ref = 1.0
        self.mu = self.mu_dim * self.H_dim / self.D_dim

    def get_mesh_generator_params(self):
        has = self.sulci_n > 0
        return {'width': self.L, 'height': self.H,
                'sulcus_depth': self.sulci_h if has else 0, 'sulcus_width': self.sulci_w if has else 0,
                'mesh_size': self.mesh_size, 'refinement_factor': self.refinement_factor, 'output_dir': None}

    def to_dict(self):
        """Serialisable view (the reference's own to_dict raises NameError; this one works)."""
        def mu_repr(m):
            if isinstance(m, StepUptakeOpen):
                return {'type': 'StepUptakeOpen', 'mu_base': m.mu_base, 'mu_open': m.mu_open, 'sulcus_left_x': m.xL,
                        'sulcus_right_x': m.xR, 'L_c': m.L_c, 'Gamma': m.Gamma}
            return m
        out = {'mode': self.mode,
               'dimensional': {k: getattr(self, k) for k in ('L_dim', 'H_dim', 'sulci_n', 'sulci_h_dim', 'sulci_w_dim',
                                                             'mesh_size_dim', 'refinement_factor', 'U_ref_dim', 'D_dim')}}
        out['dimensional']['mu_dim'] = mu_repr(self.mu_dim)
        if hasattr(self, 'L_ref'):
            out['non_dimensional'] = {k: getattr(self, k) for k in ('L_ref', 'L', 'H', 'sulci_h', 'sulci_w', 'mesh_size', 'U_ref', 'D')}
            out['non_dimensional']['mu'] = mu_repr(self.mu)
        out['computed_metrics'] = {k: getattr(self, k) for k in ('Pe', 'Re') if getattr(self, k, None) is not None}
        return out

    @classmethod
    def from_dict(cls, params_dict):
        dims = {k: v for k, v in params_dict.get('dimensional', {}).items() if k != 'mu_dim'}
        return cls(mode=params_dict.get('mode', 'adv-diff'), **dims)

    def __str__(self):
        head = f"Simulation Parameters ({self.mode.title()} Mode):"
        body = [f"  Domain: L={self.L_dim}×H={self.H_dim}mm",
                f"  Mesh: size={self.mesh_size_dim}mm, refinement={self.refinement_factor}×",
                f"  Sulci: n={self.sulci_n}, {self.sulci_w_dim}×{self.sulci_h_dim}mm"]
        return '\n'.join([head] + body)


# (width, depth, key, aspect-ratio category) of the reference's systematic catalogue, parameters.py:365-402
_CATALOGUE = [
    (1.0, 0.2, 'very_wide_tiny', 'very_wide'), (1.0, 0.3, 'very_wide_medium', 'very_wide'), (1.0, 0.5, 'very_wide_large', 'very_wide'),
    (0.5, 0.3, 'mod_wide_small', 'mod_wide'), (0.8, 0.6, 'mod_wide_medium', 'mod_wide'), (1.0, 0.9, 'mod_wide_large', 'mod_wide'),
    (0.2, 0.2, 'square_small', 'square'), (0.5, 0.5, 'square_medium', 'square'), (0.7, 0.7, 'square_large', 'square'),
    (0.5, 0.8, 'mod_deep_small', 'mod_deep'), (0.5, 1.0, 'reference', 'mod_deep'), (1.0, 1.5, 'mod_deep_large', 'mod_deep'),
    (0.3, 1.0, 'deep_small', 'deep'), (0.5, 1.5, 'deep_medium', 'deep'), (0.4, 2.0, 'deep_large', 'deep'),
    (0.25, 1.5, 'very_deep_small', 'very_deep'), (0.15, 1.8, 'very_deep_large', 'very_deep'), (0.1, 2.0, 'very_deep_extreme', 'very_deep'),
    (1.0, 0.05, 'micro_depth_wide', 'special'), (0.05, 1.0, 'micro_width_deep', 'special'), (1.0, 2.0, 'largest', 'special'),
    (0.01, 0.01, 'micro_square', 'special'), (1.0, 1.0, 'macro_square', 'special'),
]
_SMALL_PANEL = [
    (0.03, 0.03, 'small_sq_030', 'small'), (0.05, 0.05, 'small_sq_050', 'small'), (0.08, 0.08, 'small_sq_080', 'small'),
    (0.10, 0.10, 'small_sq_100', 'small'), (0.10, 0.05, 'small_wide_100x050', 'small'), (0.05, 0.10, 'small_deep_050x100', 'small'),
]


def create_geometry_variations(base_params, max_width=1.0, small_thresh=0.10, include_small=False):
    """Sulcus (width, depth) catalogue used by the geometry sweeps; same keys as the reference."""
    H, L = float(base_params.H_dim), float(base_params.L_dim)
    items = list(_CATALOGUE) + (list(_SMALL_PANEL) if include_small else [])
    configs = {}
    for width, depth, key, category in items:
        w = min(width, max_width)
        ar = depth / w if w > 0 else float('inf')
        rel = max(w / H, depth / H)
        configs[key] = {
            'L_dim': base_params.L_dim, 'H_dim': base_params.H_dim, 'mode': base_params.mode,
            'sulci_w_dim': w, 'sulci_h_dim': depth,
            'name': f"{key} ({w:.2f}x{depth:.2f} mm, AR={ar:.2f})",
            'aspect_ratio': ar, 'aspect_ratio_category': category,
            'width_ratio_L': w / L, 'width_over_H': w / H, 'depth_over_H': depth / H, 'depth_ratio': depth / H,
            'is_small': bool(rel <= small_thresh),
            'smallness_reason': f"max(w/H, h/H) = {rel:.3f} {'<= ' if rel <= small_thresh else '> '} {small_thresh:.2f}",
            'small_threshold': small_thresh,
        }
    return configs
