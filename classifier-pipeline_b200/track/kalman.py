"""Constant-velocity Kalman filter over a track's centroid (track/kalman.py:5-26).

The reference uses ``cv2.KalmanFilter(4, 2)`` in float32 with identity measurement noise,
process noise 0.03 I and zero initial state / covariance.  Host-side bookkeeping by design
(BASELINE north_star): cv2's filter is used when OpenCV is importable so that the float32
rounding is identical; otherwise the same recursion runs in numpy float32.
"""
import numpy as np

try:  # pragma: no cover - depends on the image
    import cv2 as _cv2
except Exception:  # pragma: no cover
    _cv2 = None

_F = np.array([[1, 0, 1, 0], [0, 1, 0, 1], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
_H = np.eye(2, 4, dtype=np.float32)
_Q = np.eye(4, dtype=np.float32) * np.float32(0.03)
_R = np.eye(2, dtype=np.float32)


class _NumpyKalman:
    def __init__(self):
        self.state = np.zeros((4, 1), np.float32)
        self.cov = np.zeros((4, 4), np.float32)
        self.state_pre = self.state
        self.cov_pre = self.cov

    def predict(self):
        self.state_pre = _F @ self.state
        self.cov_pre = (_F @ self.cov) @ _F.T + _Q
        self.state = self.state_pre.copy()
        self.cov = self.cov_pre.copy()
        return self.state_pre

    def correct(self, z):
        z = np.asarray(z, np.float32).reshape(2, 1)
        hp = _H @ self.cov_pre
        s = hp @ _H.T + _R
        gain = np.linalg.solve(s.astype(np.float64), hp.astype(np.float64)).T.astype(np.float32)
        self.state = self.state_pre + gain @ (z - _H @ self.state_pre)
        self.cov = self.cov_pre - gain @ hp
        return self.state


class Kalman:
    def __init__(self, use_cv2=None):
        self.use_cv2 = (_cv2 is not None) if use_cv2 is None else use_cv2
        self.reset_kalman()

    def reset_kalman(self):
        if self.use_cv2:
            k = _cv2.KalmanFilter(4, 2)
            k.measurementMatrix = _H.copy()
            k.transitionMatrix = _F.copy()
            k.processNoiseCov = _Q.copy()
            self.kalman = k
        else:
            self.kalman = _NumpyKalman()

    def predict(self):
        return self.kalman.predict()

    def correct(self, rect):
        self.kalman.correct(np.array([np.float32(rect.centroid[0]), np.float32(rect.centroid[1])], np.float32))
