"""Tracking configuration *data* for the hot path.

The reference's config package (src/config/, 1.8 kLoC of YAML/TOML plumbing) is out of scope
(SURVEY.md section 2 row 17); the extractor only reads ``TrackingConfig`` fields, so this
module carries the same field names and default values (config/trackingconfig.py:125-204,
config/trackingmotionconfig.py:24-59) as plain classes.  Objects of the reference's own
``TrackingConfig`` are accepted everywhere by duck typing.
"""
import copy


class ThresholdConfig:
    def __init__(self, camera_model, temp_thresh, background_thresh, default=False, min_temp_thresh=None,
                 max_temp_thresh=None, track_min_delta=1.0, track_max_delta=150):
        self.camera_model = camera_model
        self.temp_thresh = temp_thresh
        self.background_thresh = background_thresh
        self.default = default
        self.min_temp_thresh = min_temp_thresh
        self.max_temp_thresh = max_temp_thresh
        self.track_min_delta = track_min_delta
        self.track_max_delta = track_max_delta

    def as_dict(self):
        return dict(vars(self))


class TrackingMotionConfig:
    def __init__(self, camera_thresholds, dynamic_thresh=True):
        self.camera_thresholds = camera_thresholds
        self.dynamic_thresh = dynamic_thresh

    @classmethod
    def get_defaults(cls):
        return cls({
            "lepton3": ThresholdConfig("lepton3", 2900, 20, default=True),
            "lepton3.5": ThresholdConfig("lepton3.5", 28000, 50),
            "IR": ThresholdConfig("IR", None, 12),
        })

    def threshold_for_model(self, camera_model):
        """Exact model, else the entry flagged default (trackingmotionconfig.py:76-87)."""
        if self.camera_thresholds is None:
            return None
        found = self.camera_thresholds.get(camera_model)
        if found:
            return found
        for candidate in self.camera_thresholds.values():
            if candidate.default:
                return candidate
        return self.camera_thresholds["default-model"]

    def as_dict(self):
        return {"camera_thresholds": {k: v.as_dict() for k, v in self.camera_thresholds.items()},
                "dynamic_thresh": self.dynamic_thresh}


class TrackingConfig:
    """Field-for-field the reference's ``TrackingConfig`` defaults for one tracker type."""

    def __init__(self, type="thermal"):
        self.tracker = "RegionTracker"
        self.type = type
        self.motion = TrackingMotionConfig.get_defaults()
        self.edge_pixels = 1
        self.frame_padding = 4
        self.min_dimension = 0
        self.track_smoothing = False
        self.denoise = True
        self.high_quality_optical_flow = False
        self.max_tracks = None
        self.filters = {"track_overlap_ratio": 0.5, "min_duration_secs": 0, "track_min_offset": 4.0,
                        "track_min_mass": 2.0, "moving_vel_thresh": 4}
        self.areas_of_interest = {"min_mass": 4.0, "pixel_variance": 2.0, "cropped_regions_strategy": "cautious"}
        self.aoi_min_mass = 4.0
        self.aoi_pixel_variance = 2.0
        self.cropped_regions_strategy = "cautious"
        self.track_min_offset = 4.0
        self.track_min_mass = 2.0
        self.track_overlap_ratio = 0.5
        self.min_duration_secs = 0
        self.min_tag_confidence = 0.8
        self.enable_track_output = True
        self.moving_vel_thresh = 4
        self.min_moving_frames = 2
        self.max_blank_percent = 30
        self.max_mass_std_percent = 0.55
        self.max_jitter = 20
        self.params = {"base_distance_change": 450, "min_mass_change": 20, "restrict_mass_after": 1.5,
                       "mass_change_percent": 0.55, "max_distance": 2000, "max_blanks": 18,
                       "velocity_multiplier": 2, "base_velocity": 2}
        self.filter_regions_pre_match = True
        self.min_hist_diff = None
        self.verbose = False
        if type == "IR":
            self.filters["min_duration_secs"] = 0
            self.filter_regions_pre_match = False
            self.areas_of_interest["pixel_variance"] = 0
            self.areas_of_interest["min_mass"] = 0
            self.filters["track_min_offset"] = 7
            self.track_min_offset = 20
            self.min_dimension = 10
            self.frame_padding = 10
            self.edge_pixels = 0
            self.params = {"base_distance_change": 12000, "min_mass_change": None, "restrict_mass_after": 1.5,
                           "mass_change_percent": None, "max_distance": 30752, "max_blanks": 18,
                           "velocity_multiplier": 8, "base_velocity": 10}

    @classmethod
    def get_defaults(cls):
        return {"thermal": cls("thermal"), "IR": cls("IR")}

    @classmethod
    def get_type_defaults(cls, type):
        return cls(type)

    def get(self, type):
        """A single-type config used where the reference passes the ``{type: config}`` mapping."""
        return self if type == self.type else None

    def as_dict(self):
        d = copy.deepcopy({k: v for k, v in vars(self).items() if k not in ("motion", "verbose")})
        d["motion"] = self.motion.as_dict()
        return d


class Config:
    """The slice of ``config.Config`` the extraction path reads."""

    def __init__(self):
        self.tracking = TrackingConfig.get_defaults()
        self.use_opt_flow = False
        self.worker_threads = 0

    @classmethod
    def get_defaults(cls):
        return cls()
