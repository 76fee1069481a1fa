// kmpc_map.inl -- occupancy map -> packed circles (host code, part of libkmpc.so; included by kmpc.cu).
//
// Replaces the stand-alone script obstacle_handling/static_obstacle.py:12-56 (threshold the map at 127, distance transform of the
// occupied region, then greedily: take the largest inscribed circle, blank its disc, repeat until the largest remaining distance
// is below MIN_RADIUS) and makes its result usable: the script only paints the circles into an image; here they come back as
// centres + radii, the candidate set kmpc_select_obstacles / kmpc_environment_loop filter per agent (environment.py:48-56).
// One-off preprocessing of a map (1522 x 817 pixels for rrc_lab.pgm), not the hot path: plain C++ on the host.
//
// The script's arithmetic is OpenCV's, so three of its routines are restated to the bit:
//   cv2.threshold(img, 127, 255, THRESH_BINARY) + bitwise_not      (:23, :32)   occupied = pixel <= 127
//   cv2.distanceTransform(occupied, DIST_L2, 5)                    (:35)        two-pass 5x5 chamfer transform in float32 with the
//                                                                               weights 1, 1.4, 2.1969 (NOT the exact Euclidean distance)
//   cv2.minMaxLoc / cv2.circle(dist, center, int(maxVal), 0, -1)   (:40, :57)   first maximum in raster order; filled midpoint circle
#include <algorithm>
#include <vector>

namespace kmpc_map {

static void chamfer5(const unsigned char *occ, int w, int h, std::vector<float> &dist) {
    // OpenCV 4.x runs the 5x5 chamfer passes in float32 (weights 1, 1.4f, 2.1969f; every sum rounded to float), so do we
    const int B = 2, step = w + 2 * B;
    const float HV = 1.0f, DG = 1.4f, LG = 2.1969f;
    const float DMAX = 3.4028235e38f / 2;
    std::vector<float> tmp((size_t)step * (h + 2 * B), DMAX);
    for (int i = 0; i < h; ++i) {
        float *t = tmp.data() + (size_t)(i + B) * step + B;
        const unsigned char *s = occ + (size_t)i * w;
        for (int j = 0; j < w; ++j) {
            if (!s[j]) { t[j] = 0; continue; }
            float t0 = t[j - 2 * step - 1] + LG, v;
            v = t[j - 2 * step + 1] + LG; if (t0 > v) t0 = v;
            v = t[j - step - 2] + LG; if (t0 > v) t0 = v;
            v = t[j - step - 1] + DG; if (t0 > v) t0 = v;
            v = t[j - step] + HV; if (t0 > v) t0 = v;
            v = t[j - step + 1] + DG; if (t0 > v) t0 = v;
            v = t[j - step + 2] + LG; if (t0 > v) t0 = v;
            v = t[j - 1] + HV; if (t0 > v) t0 = v;
            t[j] = t0 > DMAX ? DMAX : t0;
        }
    }
    dist.resize((size_t)w * h);
    for (int i = h - 1; i >= 0; --i) {
        float *t = tmp.data() + (size_t)(i + B) * step + B;
        float *d = dist.data() + (size_t)i * w;
        for (int j = w - 1; j >= 0; --j) {
            float t0 = t[j], v;
            if (t0 > HV) {
                v = t[j + 2 * step + 1] + LG; if (t0 > v) t0 = v;
                v = t[j + 2 * step - 1] + LG; if (t0 > v) t0 = v;
                v = t[j + step + 2] + LG; if (t0 > v) t0 = v;
                v = t[j + step + 1] + DG; if (t0 > v) t0 = v;
                v = t[j + step] + HV; if (t0 > v) t0 = v;
                v = t[j + step - 1] + DG; if (t0 > v) t0 = v;
                v = t[j + step - 2] + LG; if (t0 > v) t0 = v;
                v = t[j + 1] + HV; if (t0 > v) t0 = v;
                t[j] = t0;
            }
            d[j] = t0;
        }
    }
}

// cv2.circle(img, (cx, cy), r, 0, thickness=-1): the scan lines of OpenCV's filled midpoint circle, clipped to the image
template <class F>
static void filled_circle(int cx, int cy, int radius, int w, int h, F &&hline) {
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    auto line = [&](int y, int x0, int x1) {
        if (y < 0 || y >= h) return;
        if (x0 < 0) x0 = 0;
        if (x1 > w - 1) x1 = w - 1;
        if (x0 <= x1) hline(y, x0, x1);
    };
    while (dx >= dy) {
        line(cy - dy, cx - dx, cx + dx); line(cy + dy, cx - dx, cx + dx);
        line(cy - dx, cx - dy, cx + dy); line(cy + dx, cx - dy, cx + dy);
        dy++;
        err += plus; plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask; dx += mask; minus -= mask & 2;
    }
}

}  // namespace kmpc_map

extern "C" int kmpc_map_distance(const unsigned char *image, int w, int h, int threshold, float *dist_out) {
    if (!image || w < 1 || h < 1 || !dist_out) return KMPC_E_BADARG;
    std::vector<unsigned char> occ((size_t)w * h);
    for (size_t i = 0; i < occ.size(); ++i) occ[i] = image[i] > threshold ? 0 : 255;
    std::vector<float> dist;
    kmpc_map::chamfer5(occ.data(), w, h, dist);
    std::copy(dist.begin(), dist.end(), dist_out);
    return 0;
}

extern "C" int kmpc_map_to_circles(const unsigned char *image, int w, int h, int threshold, double min_radius, int max_circles,
                                   int32_t *centers_out, int32_t *radii_out, int32_t *count_out) {
    if (!image || w < 1 || h < 1 || max_circles < 0 || !count_out || (max_circles > 0 && (!centers_out || !radii_out))) return KMPC_E_BADARG;
    std::vector<unsigned char> occ((size_t)w * h);
    for (size_t i = 0; i < occ.size(); ++i) occ[i] = image[i] > threshold ? 0 : 255;   // threshold + bitwise_not: occupied where the map is dark
    std::vector<float> dist;
    kmpc_map::chamfer5(occ.data(), w, h, dist);
    // the script's loop takes the global maximum (first in raster order) every time; the map only ever loses values (discs are set
    // to zero), so one pass over the pixels sorted by (value descending, raster index ascending) visits the same maxima in the same order
    std::vector<int> order;
    order.reserve(occ.size());
    for (size_t i = 0; i < dist.size(); ++i) if (dist[i] >= (float)min_radius) order.push_back((int)i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return dist[a] > dist[b]; });
    int n = 0;
    for (size_t q = 0; q < order.size(); ++q) {
        const int i = order[q];
        const float v = dist[i];
        if (!(v >= (float)min_radius)) continue;        // blanked by an earlier circle
        // radius = int(maxVal), centre = maxLoc (a map without a single free pixel has no finite distance: the radius is capped at the image)
        const int cx = i % w, cy = i / w, r = v < (float)(w + h) ? (int)v : w + h;
        if (n < max_circles) { centers_out[2 * n] = cx; centers_out[2 * n + 1] = cy; radii_out[n] = r; }
        ++n;
        kmpc_map::filled_circle(cx, cy, r, w, h, [&](int y, int x0, int x1) { for (int x = x0; x <= x1; ++x) dist[(size_t)y * w + x] = 0.f; });
    }
    *count_out = n;
    return 0;
}
