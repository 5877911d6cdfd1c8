"""Pulse shapes with the semantics of ``QuantumControl.Shapes`` (used by the reference's tests at
``test/test_tls_optimization.jl:12``)."""
import math

__all__ = ["blackman", "flattop", "box"]


def blackman(t, t0, T, a=0.16):
    """Blackman window between ``t0`` and ``T`` (0 outside)."""
    if not (t0 <= t <= T):
        return 0.0
    phase = 2.0 * math.pi * (t - t0) / (T - t0)
    return 0.5 * (1.0 - a - math.cos(phase) + a * math.cos(2.0 * phase))


def box(t, t0, T):
    return 1.0 if t0 <= t <= T else 0.0


def flattop(t, *, T, t_rise, t0=0.0, t_fall=None, func="blackman"):
    """1 in the middle of ``[t0, T]``, smooth switch-on over ``t_rise`` and switch-off over ``t_fall``."""
    t_fall = t_rise if t_fall is None else t_fall
    if t <= t0 or t >= T:
        return 0.0
    rising = t <= t0 + t_rise
    falling = t >= T - t_fall
    if not (rising or falling):
        return 1.0
    if func == "blackman":
        return blackman(t, t0, t0 + 2.0 * t_rise) if rising else blackman(t, T - 2.0 * t_fall, T)
    if func == "sinsq":
        x = (t - t0) / t_rise if rising else (t - T) / t_fall
        return math.sin(0.5 * math.pi * x) ** 2
    raise ValueError(f"Unknown func={func!r}. Accepted values are 'blackman' and 'sinsq'.")
