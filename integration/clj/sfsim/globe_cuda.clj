;; Host-side shim for sfsim: the tile loop of sfsim.globe/make-cube-map (globe.clj:29-80) on the GPU, through
;; libsfsim_atmosphere.so (include/sfsim_cubemap.h) and coffi / the JDK Foreign Function & Memory API, in the style of
;; src/clj/sfsim/jolt.clj.  The library fills the five arrays of every tile; reading the map tiles and encoding the
;; results (spit-jpg, spit-bytes-gz, spit-floats-gz, spit-normals) stay here.  NOT executed in this repository's CI:
;; the build image has no JVM; sfsim_b200/cubemap.py binds the same symbols and is what the tests drive.

(ns sfsim.globe-cuda
  "Cube map tiles for color, water, elevation, and normals computed through libsfsim_atmosphere.so"
  (:require
    [coffi.ffi :refer (defcfn) :as ffi]
    [coffi.mem :as mem]
    [sfsim.image :refer (slurp-image spit-jpg spit-normals)]
    [sfsim.util :refer (cube-dir cube-path index->face slurp-shorts spit-bytes-gz spit-floats-gz tile-path)])
  (:import
    (java.io
      File)))


(ffi/load-library "libsfsim_atmosphere.so")


;; the constants of make-cube-map (globe.clj:32-40) in the layout of sfsim_cubemap_config
(def config-struct
  [::mem/struct
   [[:in-level ::mem/int] [:out-level ::mem/int] [:width ::mem/int] [:surface-tilesize ::mem/int]
    [:sublevel ::mem/int] [:max-surface-level ::mem/int] [:max-color-level ::mem/int] [:padding ::mem/int]
    [:radius ::mem/double]]])


(defcfn last-error "Message of the last failed call" atmlut_last_error [] ::mem/c-string)
(defcfn world-create- "Rasters in device memory" sfsim_cubemap_world_create [::mem/int ::mem/pointer] ::mem/int)
(defcfn world-destroy "Release the rasters" sfsim_cubemap_world_destroy [::mem/pointer] ::mem/void)
(defcfn set-elevation-tile- "Upload one elevation tile" sfsim_cubemap_world_set_elevation_tile
  [::mem/pointer ::mem/int ::mem/int ::mem/int ::mem/pointer] ::mem/int)
(defcfn set-color-tile- "Upload one day (0) or night (1) tile" sfsim_cubemap_world_set_color_tile
  [::mem/pointer ::mem/int ::mem/int ::mem/int ::mem/int ::mem/pointer] ::mem/int)
(defcfn level- "All tiles of one output level, each handed to the callback" sfsim_cubemap_level
  [::mem/pointer ::mem/pointer ::mem/int ::mem/int ::mem/int ::mem/int
   [::ffi/fn [::mem/pointer ::mem/int ::mem/int ::mem/int ::mem/pointer ::mem/pointer ::mem/pointer ::mem/pointer
              ::mem/pointer ::mem/pointer] ::mem/int]
   ::mem/pointer]
  ::mem/int)


(def all-but-normal-bytes
  "SFSIM_CUBEMAP_DAY | NIGHT | WATER | SURFACE | NORMALS: spit-normals encodes the float normals itself"
  31)


(defn- check
  [status]
  (when-not (zero? status)
    (throw (RuntimeException. (str "libsfsim_atmosphere: " (last-error))))))


(defn- upload-level
  "Read every map tile of a level (cubemap.clj:232-248 world-map-tile / elevation-tile) and hand it to the library"
  [world kind level width]
  (let [n (bit-shift-left 1 level)]
    (doseq [ty (range (* 2 n)) tx (range (* 4 n))]
      (with-open [arena (mem/confined-arena)]
        (case kind
          :elevation (let [data (slurp-shorts (tile-path "tmp/elevation" level ty tx ".raw"))]
                       (check (set-elevation-tile- world level ty tx
                                                   (mem/serialize (vec data) [::mem/array ::mem/short (* width width)] arena))))
          (let [prefix (if (= kind :day) "tmp/day" "tmp/night")
                data   (:sfsim.image/data (slurp-image (tile-path prefix level ty tx ".png")))]
            (check (set-color-tile- world (if (= kind :day) 0 1) level ty tx
                                    (mem/serialize (vec data) [::mem/array ::mem/byte (* 4 width width)] arena)))))))))


(defn- bytes-of [pointer n] (byte-array (mem/deserialize-from (mem/reinterpret pointer n) [::mem/array ::mem/byte n])))
(defn- floats-of [pointer n] (float-array (mem/deserialize-from (mem/reinterpret pointer (* 4 n)) [::mem/array ::mem/float n])))


(defn make-cube-map
  "Drop-in for sfsim.globe/make-cube-map (globe.clj:29-80): same files under data/globe, pixels computed on the GPU"
  [in-level out-level]
  (let [width 675 surface-tilesize 65 sublevel 1 max-surface-level 4 max-color-level 5
        color-tilesize (inc (* (bit-shift-left 1 sublevel) (dec surface-tilesize)))
        pitch          (bit-and (+ color-tilesize 3) (bit-not 3))
        clamp          (fn [level hi] (max 0 (min hi level)))
        world*         (mem/alloc-instance ::mem/pointer)]
    (check (world-create- width world*))
    (let [world (mem/deserialize-from world* ::mem/pointer)]
      (try
        (doseq [level (distinct [(clamp in-level max-surface-level) (clamp (+ in-level sublevel) max-surface-level)])]
          (upload-level world :elevation level width))
        (upload-level world :day (clamp (+ in-level sublevel) max-color-level) width)
        (upload-level world :night (clamp (+ in-level sublevel) max-color-level) width)
        (with-open [arena (mem/confined-arena)]
          (let [config* (mem/serialize {:in-level in-level :out-level out-level :width width
                                        :surface-tilesize surface-tilesize :sublevel sublevel
                                        :max-surface-level max-surface-level :max-color-level max-color-level
                                        :padding 0 :radius 6378000.0}
                                       config-struct arena)]
            (check
              (level- world config* 0 1 256 all-but-normal-bytes
                      (fn [_ k b a day night water surface normals _normal-bytes]
                        (let [face  (index->face k)
                              image (fn [p] {:sfsim.image/width color-tilesize :sfsim.image/height color-tilesize
                                             :sfsim.image/channels 4
                                             :sfsim.image/data (bytes-of p (* 4 color-tilesize color-tilesize))})]
                          (.mkdirs (File. ^String (cube-dir "data/globe" face out-level a)))
                          (spit-jpg (cube-path "data/globe" face out-level b a ".jpg") (image day))
                          (spit-jpg (cube-path "data/globe" face out-level b a ".night.jpg") (image night))
                          (spit-bytes-gz (cube-path "data/globe" face out-level b a ".water.gz")
                                         (bytes-of water (* pitch color-tilesize)))
                          (spit-floats-gz (cube-path "data/globe" face out-level b a ".surf.gz")
                                          (floats-of surface (* 3 surface-tilesize surface-tilesize)))
                          (spit-normals (cube-path "data/globe" face out-level b a ".png")
                                        {:sfsim.image/width color-tilesize :sfsim.image/height color-tilesize
                                         :sfsim.image/data (floats-of normals (* 3 color-tilesize color-tilesize))})
                          0))
                      mem/null))))
        (finally
          (world-destroy world))))))
