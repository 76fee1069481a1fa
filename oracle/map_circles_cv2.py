"""The reference's map -> circles script as a function -- TEST INFRASTRUCTURE ONLY.

obstacle_handling/static_obstacle.py:12-56 run with its own OpenCV calls (cv2 is the script's only dependency and is present in this
image), returning the circles it draws instead of painting them:
    _, binary = cv2.threshold(image, 127, 255, cv2.THRESH_BINARY)            (:23)
    dist = cv2.distanceTransform(cv2.bitwise_not(binary), cv2.DIST_L2, 5)    (:32-35)
    loop: _, maxVal, _, maxLoc = cv2.minMaxLoc(dist); stop if maxVal < MIN_RADIUS;
          radius = int(maxVal); centre = maxLoc; cv2.circle(dist, centre, radius, 0, -1)     (:38-57)"""
import numpy as np


def circles_cv2(image, min_radius=1, limit=None):
    import cv2
    _, binary = cv2.threshold(image, 127, 255, cv2.THRESH_BINARY)
    dist = cv2.distanceTransform(cv2.bitwise_not(binary), cv2.DIST_L2, 5)
    cen, rad = [], []
    while limit is None or len(cen) < limit:
        _, max_val, _, max_loc = cv2.minMaxLoc(dist)
        if max_val < min_radius:
            break
        r = int(max_val)
        cen.append(max_loc); rad.append(r)
        cv2.circle(dist, max_loc, r, 0, -1)
    return np.array(cen, np.int32).reshape(-1, 2), np.array(rad, np.int32), dist
