"""``Model``: the user-facing variational circuit (host side).

Same constructor, properties and ``__call__`` contract as the reference's
``qml_essentials/model.py`` (ctor model.py:26-45, call model.py:1512-1737).  The
circuit program (state-prep noise -> state preparation -> L x (ansatz, encoding)
-> extra ansatz layer under data re-uploading -> end-of-circuit channels,
model.py:818-963) is recorded once per call signature with affine proxies and
executed by the CUDA backend.

B200-specific change: the (inputs x params x pulse) product batch is NOT
materialised with ``repeat`` (model.py:1452-1481).  ``_forward`` hands the three
factors to ``Script.execute`` as :class:`~.script.BatchAxis` entries and the
kernels index ``(b // div) % mod`` in place, so HBM holds ``B_I*F + B_P*L'*P``
numbers instead of ``B*(F + L'*P)``.
"""

from __future__ import annotations

import logging
import warnings
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np

from . import jaqsi as js
from . import operations as op
from . import rng
from .ansaetze import Ansaetze, Circuit, Encoding
from .gates import Gates, PulseInformation as pinfo
from .operations import KrausChannel
from .rng import safe_random_split
from .script import BatchAxis, LazyKeys
from .tape import recording

log = logging.getLogger(__name__)

_NOISE_DEFAULTS = {
    "BitFlip": 0.0,
    "PhaseFlip": 0.0,
    "Depolarizing": 0.0,
    "MultiQubitDepolarizing": 0.0,
    "AmplitudeDamping": 0.0,
    "PhaseDamping": 0.0,
    "GateError": 0.0,
    "ThermalRelaxation": None,
    "StatePreparation": 0.0,
    "Measurement": 0.0,
}


class Model:
    """A quantum circuit model."""

    def __init__(
        self,
        n_qubits: int,
        n_layers: int,
        circuit_type: Union[str, Circuit] = "No_Ansatz",
        data_reupload: Union[bool, List[List[bool]], List[List[List[bool]]]] = True,
        state_preparation: Union[str, Callable, List[Union[str, Callable]], None] = None,
        encoding: Union[Encoding, str, Callable, List[Union[str, Callable]]] = Gates.RX,
        trainable_frequencies: bool = False,
        initialization: str = "random",
        initialization_domain: List[float] = [0, 2 * np.pi],
        output_qubit: Union[List[int], int] = -1,
        shots: Optional[int] = None,
        random_seed: int = 1000,
        remove_zero_encoding: bool = True,
        repeat_batch_axis: List[bool] = [True, True, True],
        pulse_shape: str = "gaussian",
        precision: Optional[str] = None,
    ) -> None:
        """See the reference for the meaning of every argument (model.py:46-102).
        ``precision`` (``"complex64"``/``"complex128"``/None = process default) is
        the only addition."""
        self.n_qubits: int = n_qubits
        self.output_qubit = output_qubit
        self.n_layers: int = n_layers
        self.noise_params = None
        self.shots = shots
        self.remove_zero_encoding = remove_zero_encoding
        self.trainable_frequencies: bool = trainable_frequencies
        self.execution_type = "expval"
        self.repeat_batch_axis: List[bool] = list(repeat_batch_axis)

        pinfo.set_envelope(pulse_shape)

        try:
            self._sp = Gates.parse_gates(state_preparation, Gates)
        except ValueError as e:
            raise ValueError(f"Error parsing encodings: {e}")
        self.sp_pulse_params = [None for _ in self._sp]

        self._enc = encoding if isinstance(encoding, Encoding) else Encoding("hamming", encoding)
        if self._enc.is_golomb:
            self._enc._n_qubits = n_qubits
        self.n_input_feat = len(self._enc)

        # trainable frequencies, initialised as in arXiv:2309.03279v2 (model.py:150)
        self.enc_params = np.ones((self.n_layers, self.n_qubits, self.n_input_feat))
        self._zero_inputs = False

        self.data_reupload = data_reupload  # also sets degree / frequencies / has_dru
        impl_n_layers = n_layers + 1 if self.has_dru else n_layers  # model.py:163-166

        if isinstance(circuit_type, str):
            self.pqc = getattr(Ansaetze, circuit_type or "No_Ansatz")()
        else:
            self.pqc = circuit_type()

        self._params_shape: Tuple[int, int] = (
            impl_n_layers, self.pqc.n_params_per_layer(self.n_qubits))
        try:
            n_pulse = self.pqc.n_pulse_params_per_layer(self.n_qubits)
        except NotImplementedError:
            n_pulse = 0
        self._pulse_params_shape: Tuple[int, int] = (impl_n_layers, n_pulse)

        self._batch_shape = None
        self._inialization_strategy = initialization
        self._initialization_domain = initialization_domain
        self.random_key = self.initialize_params(rng.key(random_seed))
        self.pulse_params = np.ones((1, *self._pulse_params_shape))

        self.script = js.Script(f=self._variational, n_qubits=self.n_qubits,
                                precision=precision)

    # ------------------------------------------------------------------ noise
    @property
    def noise_params(self):
        return self._noise_params

    @noise_params.setter
    def noise_params(self, kvs) -> None:
        """Fill defaults, warn on unknown keys, validate ThermalRelaxation
        (model.py:223-299)."""
        if kvs is not None and all(v == 0.0 for v in kvs.values()):
            kvs = None
        if kvs is not None:
            for k, v in _NOISE_DEFAULTS.items():
                kvs.setdefault(k, v)
            for k in kvs:
                if k not in _NOISE_DEFAULTS:
                    warnings.warn(f"Noise type {k} is not supported by this package",
                                  UserWarning)
            tr = kvs["ThermalRelaxation"]
            if isinstance(tr, dict):
                for k in ("t1", "t2", "t_factor"):
                    tr.setdefault(k, 0.0)
                for k in tr:
                    if k not in ("t1", "t2", "t_factor"):
                        warnings.warn(
                            f"Thermal Relaxation parameter {k} is not supported "
                            "by this package", UserWarning)
                if not all(tr.values()) or tr["t2"] > 2 * tr["t1"]:
                    warnings.warn(
                        "Received invalid values for Thermal Relaxation noise "
                        "parameter. Thermal relaxation is not applied!", UserWarning)
                    kvs["ThermalRelaxation"] = 0.0
        self._noise_params = kvs

    # ------------------------------------------------------------- measurement
    @property
    def output_qubit(self) -> List[int]:
        return self._output_qubit

    @output_qubit.setter
    def output_qubit(self, value) -> None:
        if isinstance(value, list):
            assert len(value) <= self.n_qubits, (
                f"Size of output_qubit {len(value)} cannot be larger than "
                f"number of qubits {self.n_qubits}.")
        elif isinstance(value, (int, np.integer)):
            if value == -1:
                value = list(range(self.n_qubits))
            else:
                assert value < self.n_qubits, (
                    f"Output qubit {value} cannot be larger than {self.n_qubits}.")
                value = [int(value)]
        self._output_qubit = value

    @property
    def execution_type(self) -> str:
        return self._execution_type

    @execution_type.setter
    def execution_type(self, value: str) -> None:
        """Sets the per-element result shape (model.py:340-385)."""
        n_out = len(self.output_qubit)
        if value == "density":
            self._result_shape = (2**n_out, 2**n_out)
        elif value == "expval":
            self._result_shape = (n_out,)
        elif value == "probs":
            self._result_shape = (2,) * n_out
        elif value == "state":
            self._result_shape = (2**n_out,)
        else:
            raise ValueError(f"Invalid execution type: {value}.")
        if value == "state" and not self.all_qubit_measurement:
            warnings.warn(
                f"{value} measurement does ignore output_qubit, which is "
                f"{self.output_qubit}.", UserWarning)
        if value == "probs" and self.shots is None:
            warnings.warn("Setting execution_type to probs without specifying shots.",
                          UserWarning)
        if value == "density" and self.shots is not None:
            raise ValueError("Setting execution_type to density with shots not None.")
        self._execution_type = value

    @property
    def shots(self) -> Optional[int]:
        return self._shots

    @shots.setter
    def shots(self, value: Optional[int]) -> None:
        if type(value) is int and value <= 0:
            value = None
        self._shots = value

    # -------------------------------------------------------------- parameters
    @property
    def params(self) -> np.ndarray:
        return self._params

    @params.setter
    def params(self, value) -> None:
        value = np.asarray(value) if not isinstance(value, np.ndarray) else value
        if value.ndim == 2:
            value = value.reshape(1, *value.shape)
        self._params = value

    @property
    def enc_params(self) -> np.ndarray:
        return self._enc_params

    @enc_params.setter
    def enc_params(self, value) -> None:
        self._enc_params = value

    @property
    def pulse_params(self) -> np.ndarray:
        return self._pulse_params

    @pulse_params.setter
    def pulse_params(self, value) -> None:
        self._pulse_params = value

    @property
    def data_reupload(self) -> np.ndarray:
        return self._data_reupload

    @data_reupload.setter
    def data_reupload(self, value) -> None:
        """bool | (L, n) | (L, n, F) mask -> boolean (L, n, F); updates degree,
        frequencies and has_dru (model.py:451-512)."""
        full = (self.n_layers, self.n_qubits, self.n_input_feat)
        if isinstance(value, (bool, np.bool_)):
            mask = np.ones(full) if value else np.zeros(full)
            if not value:
                mask[0][0] = 1
        else:
            mask = np.asarray(value)
            if mask.ndim == 2:
                assert mask.shape == full[:2], (
                    f"Data reuploading array has wrong shape. Expected {full[:2]} or "
                    f"{full}, got {mask.shape}.")
                mask = np.repeat(mask[..., None], self.n_input_feat, axis=2)
            assert mask.shape == full, (
                f"Data reuploading array has wrong shape. Expected {full}, got {mask.shape}.")
        self._data_reupload = mask.astype(bool)
        counts = [int(np.count_nonzero(self._data_reupload[..., i]))
                  for i in range(self.n_input_feat)]
        self.degree = tuple(self._enc.get_n_freqs(c) for c in counts)
        self.frequencies = tuple(self._enc.get_spectrum(c) for c in counts)
        self._has_dru = bool(max(int(np.max(f)) for f in self.frequencies) > 1)

    @property
    def degree(self) -> Tuple:
        return self._degree

    @degree.setter
    def degree(self, value: Tuple):
        self._degree = value

    @property
    def frequencies(self) -> Tuple:
        return self._frequencies

    @frequencies.setter
    def frequencies(self, value: Tuple):
        self._frequencies = value

    def exact_spectrum(self, method: str = "tree"):
        raise NotImplementedError(
            "the symbolic FourierTree (coefficients.py:240-964) is outside the "
            "B200 backend scope")

    @property
    def has_dru(self) -> bool:
        return self._has_dru

    @property
    def all_qubit_measurement(self) -> bool:
        return self.output_qubit == list(range(self.n_qubits))

    @property
    def batch_shape(self) -> Tuple[int, ...]:
        """(B_I, B_P, B_R) of the last call, (1, 1, 1) before any call."""
        return (1, 1, 1) if self._batch_shape is None else self._batch_shape

    @property
    def eff_batch_shape(self) -> np.ndarray:
        shape = np.array(self.batch_shape) * self.repeat_batch_axis
        return shape[shape != 0]

    def initialize_params(self, random_key=None, repeat: int = 1,
                          initialization: Optional[str] = None,
                          initialization_domain: Optional[List[float]] = None):
        """(Re-)draw ``repeat`` parameter sets; returns the advanced key
        (model.py:631-722).  Strategies: random | zeros | pi | zero-controlled |
        pi-controlled."""
        shape = (repeat, *self._params_shape)
        initialization = initialization or self._inialization_strategy
        lo, hi = initialization_domain or self._initialization_domain
        random_key, sub_key = safe_random_split(
            random_key if random_key is not None else self.random_key)

        def draw():
            return rng.uniform(sub_key, shape, minval=lo, maxval=hi)

        def with_controls(params, value):
            idx = self.pqc.get_control_indices(self.n_qubits)
            if idx is None:
                warnings.warn(
                    f"Specified {initialization} but circuit does not contain "
                    "controlled rotation gates. Parameters are intialized randomly.",
                    UserWarning)
                return params
            params = np.array(params)
            if len(idx) == 3 and None in idx:
                params[:, :, idx[0]:idx[1]:idx[2]] = value
            else:
                params[:, :, idx] = value
            return params

        if initialization == "random":
            self.params = draw()
        elif initialization == "zeros":
            self.params = np.zeros(shape)
        elif initialization == "pi":
            self.params = np.ones(shape) * np.pi
        elif initialization == "zero-controlled":
            self.params = with_controls(draw(), 0)
        elif initialization == "pi-controlled":
            self.params = with_controls(draw(), np.pi)
        else:
            raise Exception("Invalid initialization method")
        return random_key

    def transform_input(self, inputs, enc_params):
        """Linear input scaling of arXiv:2309.03279v2 (model.py:724-744)."""
        return inputs * enc_params

    # ---------------------------------------------------------- circuit program
    def _iec(self, inputs, data_reupload, enc: Encoding, enc_params,
             noise_params=None, random_key=None) -> None:
        """Input-encoding layer (model.py:746-816)."""
        if self.remove_zero_encoding and self._zero_inputs and self.batch_shape[0] == 1:
            return
        if enc.is_golomb:
            if data_reupload[:, 0].any():
                random_key, sub_key = safe_random_split(random_key)
                scale = enc_params[:, 0].mean()  # mean over qubits (model.py:793)
                enc[0](self.transform_input(inputs[..., 0], scale),
                       wires=list(range(self.n_qubits)), noise_params=noise_params,
                       random_key=sub_key)
            return
        for q in range(self.n_qubits):
            for idx in range(inputs.shape[-1]):
                if data_reupload[q, idx]:
                    random_key, sub_key = safe_random_split(random_key)
                    enc[idx](self.transform_input(inputs[..., idx], enc_params[q, idx]),
                             wires=q, noise_params=noise_params, random_key=sub_key)

    def _variational(self, params, inputs, pulse_params=None, random_key=None,
                     enc_params=None, gate_mode: str = "unitary", noise_params=None) -> None:
        """Record the whole circuit (model.py:818-963)."""
        if len(params.shape) > 2 and params.shape[0] == 1:
            params = params[0]
        if len(inputs.shape) > 1 and inputs.shape[0] == 1:
            inputs = inputs[0]
        if enc_params is None:
            if self.trainable_frequencies:
                warnings.warn(
                    "Explicit call to `_circuit` or `_variational` detected: "
                    "`enc_params` is None, using `self.enc_params` instead.", RuntimeWarning)
            enc_params = self.enc_params
        if pulse_params is None:
            pulse_params = self.pulse_params
        if len(pulse_params.shape) > 2 and pulse_params.shape[0] == 1:
            pulse_params = pulse_params[0]
        if noise_params is None and self.noise_params is not None:
            warnings.warn(
                "Explicit call to `_circuit` or `_variational` detected: "
                "`noise_params` is None, using `self.noise_params` instead.", RuntimeWarning)
            noise_params = self.noise_params
        if noise_params is not None:
            if random_key is None:
                warnings.warn(
                    "Explicit call to `_circuit` or `_variational` detected: "
                    "`random_key` is None, using the model's key instead.", RuntimeWarning)
                random_key = self.random_key
            self._apply_state_prep_noise(noise_params=noise_params)

        for q in range(self.n_qubits):
            for sp, sp_pulse in zip(self._sp, self.sp_pulse_params):
                random_key, sub_key = safe_random_split(random_key)
                sp(wires=q, pulse_params=sp_pulse, noise_params=noise_params,
                   random_key=sub_key, gate_mode=gate_mode)

        for layer in range(self.n_layers):
            random_key, sub_key = safe_random_split(random_key)
            self.pqc(params[layer], self.n_qubits, pulse_params=pulse_params[layer],
                     noise_params=noise_params, random_key=sub_key, gate_mode=gate_mode)
            random_key, sub_key = safe_random_split(random_key)
            self._iec(inputs, data_reupload=self.data_reupload[layer], enc=self._enc,
                      enc_params=enc_params[layer], noise_params=noise_params,
                      random_key=sub_key)

        if self.has_dru:
            random_key, sub_key = safe_random_split(random_key)
            self.pqc(params[self.n_layers], self.n_qubits, pulse_params=pulse_params[-1],
                     noise_params=noise_params, random_key=sub_key, gate_mode=gate_mode)

        if noise_params is not None:
            self._apply_general_noise(noise_params=noise_params)

    def _build_obs(self) -> Tuple[str, List[op.Operation]]:
        """execution_type / output_qubit -> (measurement type, observables)
        (model.py:965-998)."""
        if self.execution_type in ("density", "state", "probs"):
            return self.execution_type, []
        if self.execution_type == "expval":
            obs: List[op.Operation] = []
            for spec in self.output_qubit:
                if isinstance(spec, (int, np.integer)):
                    obs.append(op.PauliZ(wires=int(spec), record=False))
                else:
                    obs.append(js.build_parity_observable(list(spec)))
            return "expval", obs
        raise ValueError(f"Invalid execution_type: {self.execution_type}.")

    def _apply_state_prep_noise(self, noise_params) -> None:
        """BitFlip(StatePreparation) on every qubit (model.py:1000-1020)."""
        p = noise_params.get("StatePreparation", 0.0)
        if p > 0:
            for q in range(self.n_qubits):
                op.BitFlip(p, wires=q)

    def _apply_general_noise(self, noise_params) -> None:
        """End-of-circuit channels per qubit (model.py:1022-1064)."""
        amp = noise_params.get("AmplitudeDamping", 0.0)
        phase = noise_params.get("PhaseDamping", 0.0)
        thermal = noise_params.get("ThermalRelaxation", 0.0)
        meas = noise_params.get("Measurement", 0.0)
        for q in range(self.n_qubits):
            if amp > 0:
                op.AmplitudeDamping(amp, wires=q)
            if phase > 0:
                op.PhaseDamping(phase, wires=q)
            if meas > 0:
                op.BitFlip(meas, wires=q)
            if isinstance(thermal, dict):
                tg = self._get_circuit_depth() * thermal["t_factor"]
                op.ThermalRelaxationError(1.0, thermal["t1"], thermal["t2"], tg, q)

    def _get_circuit_depth(self, inputs=None) -> int:
        """Critical-path length of the noise-free tape; Barriers count, channels do
        not (model.py:1066-1122).  Unlike the reference, ``_zero_inputs`` is restored
        afterwards (the reference leaks the flag set by ``_inputs_validation(None)``
        into the recording that follows; see DESIGN.md 'deviations')."""
        if hasattr(self, "_cached_circuit_depth"):
            return self._cached_circuit_depth
        saved_zero, saved_noise = self._zero_inputs, self._noise_params
        inputs = self._inputs_validation(inputs)
        self._noise_params = None
        try:
            with recording() as tape:
                self._variational(
                    self.params[0] if self.params.ndim == 3 else self.params,
                    inputs[0] if inputs.ndim == 2 else inputs, noise_params=None)
        finally:
            self._noise_params = saved_noise
            self._zero_inputs = saved_zero
        busy: Dict[int, int] = {}
        depth = 0
        for gate in tape:
            if isinstance(gate, KrausChannel):
                continue
            end = max((busy.get(w, 0) for w in gate.wires), default=0) + 1
            for w in gate.wires:
                busy[w] = end
            depth = max(depth, end)
        self._cached_circuit_depth = depth
        return depth

    def draw(self, inputs=None, figure: str = "text", **kwargs: Any):
        """Text listing of the noise-free circuit (model.py:1124-1180; graphical
        back ends are out of scope)."""
        inputs = self._inputs_validation(inputs)
        params = self.params[0] if self.params.ndim == 3 else self.params
        inp = inputs[0] if inputs.ndim == 2 else inputs
        saved = self._noise_params
        self._noise_params = None
        try:
            script = js.Script(f=self._variational, n_qubits=self.n_qubits)
            return script.draw(figure=figure, args=(params, inp),
                               kwargs={"noise_params": None}, **kwargs)
        finally:
            self._noise_params = saved

    def __repr__(self) -> str:
        return self.draw(figure="text")

    __str__ = __repr__

    # ------------------------------------------------------------- validation
    def _params_validation(self, params):
        if params is None:
            return self.params
        params = np.asarray(params)
        if params.ndim == 2:
            params = np.expand_dims(params, axis=0)
        self.params = params
        return params

    def _pulse_params_validation(self, pulse_params):
        if pulse_params is None:
            return self.pulse_params
        pulse_params = np.asarray(pulse_params)
        if pulse_params.ndim == 2:
            pulse_params = np.expand_dims(pulse_params, axis=0)
        self.pulse_params = pulse_params
        return pulse_params

    def _enc_params_validation(self, enc_params):
        if enc_params is None:
            enc_params = self.enc_params
        else:
            enc_params = np.asarray(enc_params)
            self.enc_params = enc_params
        if enc_params.ndim == 1 and self.n_input_feat == 1:
            enc_params = enc_params.reshape(-1, 1)
        elif enc_params.ndim == 1 and self.n_input_feat > 1:
            raise ValueError(
                f"Input dimension {self.n_input_feat} >1 but `enc_params` has shape "
                f"{enc_params.shape}")
        return enc_params

    def _inputs_validation(self, inputs) -> np.ndarray:
        """Anything -> (batch, n_input_feat) array; flags all-zero inputs
        (model.py:1330-1389)."""
        self._zero_inputs = False
        if isinstance(inputs, list):
            inputs = np.array(np.stack(inputs))
        elif isinstance(inputs, (float, int, np.floating, np.integer)):
            inputs = np.array([inputs])
        elif inputs is None:
            inputs = np.array([[0] * self.n_input_feat])
        inputs = np.asarray(inputs)
        if not inputs.any():
            self._zero_inputs = True
        if inputs.ndim <= 1:
            if self.n_input_feat == 1:
                inputs = inputs.reshape(-1, 1)
            elif inputs.shape[0] == self.n_input_feat:
                inputs = inputs.reshape(1, -1)
            else:
                inputs = inputs.reshape(-1, 1).repeat(self.n_input_feat, axis=1)
                warnings.warn(
                    f"Expected {self.n_input_feat} inputs, but {inputs.shape[0]} "
                    "was provided, replicating input for all input features.", UserWarning)
        elif inputs.shape[1] != self.n_input_feat:
            raise ValueError(
                f"Wrong number of inputs provided. Expected {self.n_input_feat} "
                f"inputs, but input has shape {inputs.shape}.")
        return inputs

    def _postprocess_res(self, result):
        if isinstance(result, list):
            result = np.stack(result)
            if result.ndim > 1:
                result = np.moveaxis(result, 0, 1)
        return result

    def _set_batch_shape(self, inputs, params, pulse_params) -> int:
        B_I = inputs.shape[0]
        B_P = 1 if 0 in params.shape else params.shape[0]
        B_R = pulse_params.shape[0]
        self._batch_shape = (B_I, B_P, B_R)  # the only place it is set (model.py:1447)
        return int(np.prod(self.eff_batch_shape))

    def _assimilate_batch(self, inputs, params, pulse_params):
        """Materialised product batch in the reference's flat order
        ``b = (i * B_P + p) * B_R + r`` (model.py:1414-1483).  Kept for API parity;
        ``_forward`` uses batch factors instead of these copies."""
        B = self._set_batch_shape(inputs, params, pulse_params)
        B_I, B_P, B_R = self._batch_shape
        rep = self.repeat_batch_axis
        if B_I > 1 and rep[0]:
            x = inputs[:, None, None, ...]
            if rep[1]:
                x = np.repeat(x, B_P, axis=1)
            if rep[2]:
                x = np.repeat(x, B_R, axis=2)
            inputs = x.reshape(B, *x.shape[3:])
        if B_P > 1 and rep[1]:
            x = params[None, :, None, ...]
            if rep[0]:
                x = np.repeat(x, B_I, axis=0)
            if rep[2]:
                x = np.repeat(x, B_R, axis=2)
            params = x.reshape(B, *x.shape[3:])
        if B_R > 1 and rep[2]:
            x = pulse_params[None, None, ...]
            if rep[0]:
                x = np.repeat(x, B_I, axis=0)
            if rep[1]:
                x = np.repeat(x, B_P, axis=1)
            pulse_params = x.reshape(B, *x.shape[3:])
        return inputs, params, pulse_params

    def _batch_axes(self, B: int):
        """``in_axes`` entries (params, inputs, pulse_params) as batch factors."""
        sizes = self.batch_shape
        rep = self.repeat_batch_axis
        eff = [s if r else 0 for s, r in zip(sizes, rep)]
        axes = []
        for j in range(3):
            if sizes[j] <= 1:
                axes.append(None)
            elif not rep[j]:
                axes.append(BatchAxis(0, 1, sizes[j], B))  # zipped with the flat batch
            else:
                div = int(np.prod([e for e in eff[j + 1:] if e != 0] or [1]))
                axes.append(BatchAxis(0, div, sizes[j], B))
        return axes[1], axes[0], axes[2]  # order of the positional args

    def _requires_density(self) -> bool:
        """model.py:1485-1510."""
        if self.execution_type == "density":
            return True
        if self.noise_params is None:
            return False
        for k, v in self.noise_params.items():
            if k == "GateError":
                continue
            if v is not None and v > 0:
                return True
        return False

    # -------------------------------------------------------------------- call
    def __call__(self, params=None, inputs=None, pulse_params=None, enc_params=None,
                 data_reupload=None, noise_params=None, execution_type: Optional[str] = None,
                 force_mean: bool = False, gate_mode: str = "unitary"):
        """Execute the circuit (model.py:1512-1570)."""
        return self._forward(params=params, inputs=inputs, pulse_params=pulse_params,
                             enc_params=enc_params, data_reupload=data_reupload,
                             noise_params=noise_params, execution_type=execution_type,
                             force_mean=force_mean, gate_mode=gate_mode)

    def device_result(self, **call_kwargs):
        """Extension for the analysis callers: same arguments as ``__call__``, but
        the raw batched result ((B, ...) in flat batch order, full register, no
        post-processing) stays on the GPU as a ``torch`` tensor."""
        return self._forward(_device=True, **call_kwargs)

    def _forward(self, params=None, inputs=None, pulse_params=None, enc_params=None,
                 data_reupload=None, noise_params=None, execution_type=None,
                 force_mean: bool = False, gate_mode: str = "unitary", _device: bool = False):
        """Validate, batch, dispatch, reshape (model.py:1572-1737)."""
        if noise_params is not None:
            self.noise_params = noise_params
        if execution_type is not None:
            self.execution_type = execution_type
        self.gate_mode = gate_mode
        if pulse_params is not None and gate_mode != "pulse":
            raise ValueError(
                "pulse_params were provided but gate_mode is not 'pulse'. "
                "Either switch gate_mode='pulse' or do not pass pulse_params.")
        if data_reupload is not None:
            self.data_reupload = data_reupload

        params = self._params_validation(params)
        pulse_params = self._pulse_params_validation(pulse_params)
        inputs = self._inputs_validation(inputs)
        enc_params = self._enc_params_validation(enc_params)

        B = self._set_batch_shape(inputs, params, pulse_params)
        self.random_key, sub_key = safe_random_split(self.random_key)
        meas_type, obs = self._build_obs()
        exec_kwargs = dict(noise_params=self.noise_params, gate_mode=self.gate_mode)
        # everything _variational reads from `self` (not from its arguments) must
        # take part in the plan cache key
        self.script.cache_salt = (
            bool(self.remove_zero_encoding and self._zero_inputs and self.batch_shape[0] == 1),
            self._data_reupload.tobytes(), self.n_layers, self.has_dru,
        )

        shot_key = None
        if self.shots is not None:
            sub_key, shot_key = safe_random_split(sub_key)

        # sub-register density / probability outputs: trace / marginalise on the GPU so that
        # only the reduced result is copied back (jaqsi.py:79-146 on the device)
        reduce_dev = None
        if (not _device and not self.all_qubit_measurement and self.shots is None
                and self.execution_type in ("density", "probs")):
            from .script import get_executor

            try:
                ex = self.script.executor or get_executor()
            except Exception:  # no device: Script.execute raises the proper error below
                ex = None
            if ex is not None and hasattr(ex, "partial_trace"):
                reduce_dev = ex
        on_device = _device or reduce_dev is not None

        if B > 1:
            ax_params, ax_inputs, ax_pulse = self._batch_axes(B)
            keys = LazyKeys(sub_key, B)
            in_axes = (ax_params, ax_inputs, ax_pulse, BatchAxis(0, 1, B, B), None)
            result = self.script.execute(
                type=meas_type, obs=obs,
                args=(params, inputs, pulse_params, keys, enc_params),
                kwargs=exec_kwargs, in_axes=in_axes, shots=self.shots, key=shot_key,
                device_result=on_device)
        else:
            result = self.script.execute(
                type=meas_type, obs=obs,
                args=(params, inputs, pulse_params, sub_key, enc_params),
                kwargs=exec_kwargs, shots=self.shots, key=shot_key, device_result=on_device)
            if on_device:
                result = result[None]
        if _device:
            return result
        if reduce_dev is not None:
            ex = reduce_dev
            if self.execution_type == "density":
                result = ex.to_host(ex.partial_trace(result, self.n_qubits, self.output_qubit))
            elif isinstance(self.output_qubit[0], (list, tuple)):
                result = np.stack([ex.to_host(ex.marginal_probs(result, self.n_qubits, list(g)))
                                   for g in self.output_qubit])
            else:
                result = ex.to_host(ex.marginal_probs(result, self.n_qubits, self.output_qubit))
            result = np.asarray(result)
            result = result.reshape((*self.eff_batch_shape, *self._result_shape)).squeeze()
            if (self.execution_type == "probs" and force_mean
                    and len(result.shape) > 0 and self._result_shape[0] > 1):
                result = result.mean(axis=-1)
            return result

        result = self._postprocess_res(result)
        if self.execution_type == "density" and not self.all_qubit_measurement:
            result = js.partial_trace(result, self.n_qubits, self.output_qubit)
        if self.execution_type == "probs" and not self.all_qubit_measurement:
            if isinstance(self.output_qubit[0], (list, tuple)):
                result = np.stack([
                    js.marginalize_probs(result, self.n_qubits, list(g))
                    for g in self.output_qubit])
            else:
                result = js.marginalize_probs(result, self.n_qubits, self.output_qubit)

        result = np.asarray(result)
        result = result.reshape((*self.eff_batch_shape, *self._result_shape)).squeeze()
        if (self.execution_type in ("expval", "probs") and force_mean
                and len(result.shape) > 0 and self._result_shape[0] > 1):
            result = result.mean(axis=-1)
        return result
