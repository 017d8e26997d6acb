"""``CPTVMotionDetector`` and ``is_affected_by_ffc`` (piclassifier/cptvmotiondetector.py:14-234)."""
from datetime import timedelta

FFC_PERIOD = timedelta(seconds=9.9)


def is_affected_by_ffc(cptv_frame):
    """A frame taken during / just after a flat-field correction (cptvmotiondetector.py:211-222).

    parity: with integer ``time_on`` / ``last_ffc_time`` (milliseconds from the CPTV decoder) the
    difference is compared with ``FFC_PERIOD.seconds`` == 9, i.e. nine *milliseconds*."""
    if hasattr(cptv_frame, "ffc_status") and cptv_frame.ffc_status in [1, 2]:
        return True
    if cptv_frame.time_on is None or cptv_frame.last_ffc_time is None:
        return False
    if isinstance(cptv_frame.time_on, int):
        return (cptv_frame.time_on - cptv_frame.last_ffc_time) < FFC_PERIOD.seconds
    return (cptv_frame.time_on - cptv_frame.last_ffc_time) < FFC_PERIOD
