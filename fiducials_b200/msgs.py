"""Field-for-field mirrors of the fiducial_msgs messages the hot path produces
(fiducial_msgs/msg/*.msg in the reference) -- the surface a ROS node publishes from the C-ABI output."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass
class Header:  # std_msgs/Header
    seq: int = 0
    stamp: Tuple[int, int] = (0, 0)
    frame_id: str = ""


@dataclass
class Fiducial:  # fiducial_msgs/msg/Fiducial.msg:3-14 (direction is never set by the reference)
    fiducial_id: int = 0
    direction: int = 0
    x0: float = 0.0
    y0: float = 0.0
    x1: float = 0.0
    y1: float = 0.0
    x2: float = 0.0
    y2: float = 0.0
    x3: float = 0.0
    y3: float = 0.0


@dataclass
class FiducialArray:  # FiducialArray.msg
    header: Header = field(default_factory=Header)
    fiducials: List[Fiducial] = field(default_factory=list)


@dataclass
class Transform:  # geometry_msgs/Transform
    translation: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    rotation: Tuple[float, float, float, float] = (0.0, 0.0, 0.0, 1.0)  # x y z w


@dataclass
class FiducialTransform:  # FiducialTransform.msg:2-6
    fiducial_id: int = 0
    transform: Transform = field(default_factory=Transform)
    image_error: float = 0.0
    object_error: float = 0.0
    fiducial_area: float = 0.0


@dataclass
class FiducialTransformArray:  # FiducialTransformArray.msg:3-5
    header: Header = field(default_factory=Header)
    image_seq: int = 0
    transforms: List[FiducialTransform] = field(default_factory=list)


@dataclass
class FiducialMapEntry:  # FiducialMapEntry.msg:2-10
    fiducial_id: int = 0
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0
    rx: float = 0.0
    ry: float = 0.0
    rz: float = 0.0


@dataclass
class FiducialMapEntryArray:  # FiducialMapEntryArray.msg
    fiducials: List[FiducialMapEntry] = field(default_factory=list)


@dataclass
class ObjectHypothesisWithPose:  # vision_msgs/ObjectHypothesisWithPose as the reference fills it (aruco_detect.cpp:463-476)
    id: int = 0
    score: float = 0.0           # exp(-2 * object_error): [0, inf) -> (0, 1]
    position: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    orientation: Tuple[float, float, float, float] = (0.0, 0.0, 0.0, 1.0)  # x y z w


@dataclass
class Detection2D:  # vision_msgs/Detection2D: the reference only fills results[0]
    results: List[ObjectHypothesisWithPose] = field(default_factory=list)


@dataclass
class Detection2DArray:  # published instead of FiducialTransformArray when the vis_msgs parameter is set (:403,:534,:666)
    header: Header = field(default_factory=Header)
    detections: List[Detection2D] = field(default_factory=list)
