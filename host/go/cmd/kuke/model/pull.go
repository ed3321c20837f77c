// NOT COMPILED HERE (no Go toolchain) — reviewed source; see host/go/README.md.

package model

import (
	"fmt"
	"strconv"
	"strings"

	getshared "github.com/eminwux/kukeon/cmd/kuke/get/shared"
	"github.com/eminwux/kukeon/internal/gpupool"
	"github.com/spf13/cobra"
)

// NewPullCmd builds `kuke model pull <path>`: resolve a local checkpoint (directory with model.safetensors.index.json / *.safetensors /
// *.gguf, or one file) and print its tensor index.  "Pull" is local-path only; registry references are rejected by validateModels.
func NewPullCmd() *cobra.Command {
	cmd := &cobra.Command{
		Use:           "pull <path>",
		Short:         "Index a local safetensors / GGUF checkpoint (no device is touched)",
		Args:          cobra.ExactArgs(1),
		SilenceUsage:  true,
		SilenceErrors: false,
		RunE: func(cmd *cobra.Command, args []string) error {
			outputFormat, err := getshared.ParseOutputFormat(cmd)
			if err != nil {
				return err
			}
			recs, err := resolveIndexer(cmd).Index(strings.TrimSpace(args[0]))
			if err != nil {
				return err
			}
			switch outputFormat {
			case getshared.OutputFormatJSON:
				return getshared.PrintJSON(recs)
			case getshared.OutputFormatYAML:
				return getshared.PrintYAML(recs)
			default:
				if len(recs) == 0 {
					cmd.Printf("No tensors found in %q.\n", args[0])
					return nil
				}
				var total uint64
				rows := make([][]string, 0, len(recs))
				for _, r := range recs {
					total += r.NBytes
					rows = append(rows, []string{r.Name, dtypeName(r.Dtype), shapeString(r.Shape), strconv.Itoa(int(r.Shard)), formatSize(int64(r.NBytes))})
				}
				getshared.PrintTable(cmd, []string{"NAME", "DTYPE", "SHAPE", "SHARD", "SIZE"}, rows)
				cmd.Printf("%d tensors, %s\n", len(recs), formatSize(int64(total)))
				return nil
			}
		},
	}
	cmd.Flags().StringP("output", "o", "", "Output format (yaml, json, table). Default: table")
	return cmd
}

func shapeString(shape []uint64) string {
	if len(shape) == 0 {
		return "scalar"
	}
	parts := make([]string, len(shape))
	for i, s := range shape {
		parts[i] = strconv.FormatUint(s, 10)
	}
	return strings.Join(parts, "x")
}

// dtypeName spells a kk_dtype the way the manifest does (include/kukeon_gpuload.h kk_dtype).
func dtypeName(dt uint32) string {
	names := map[uint32]string{
		0: "BOOL", 1: "F4", 2: "F6_E2M3", 3: "F6_E3M2", 4: "U8", 5: "I8", 6: "F8_E5M2", 7: "F8_E4M3", 8: "F8_E8M0", 9: "I16", 10: "U16", 11: "F16",
		12: "BF16", 13: "I32", 14: "U32", 15: "F32", 16: "C64", 17: "F64", 18: "I64", 19: "U64", 32: "Q4_0", 33: "Q4_1", 34: "Q5_0", 35: "Q5_1",
		36: "Q8_0", 37: "Q2_K", 38: "Q3_K", 39: "Q4_K", 40: "Q5_K", 41: "Q6_K", 42: "Q8_K", 43: "IQ4_NL", 44: "IQ4_XS", 45: "MXFP4", 46: "IQ2_XXS",
		47: "IQ2_XS", 48: "IQ2_S", 49: "IQ3_XXS", 50: "IQ3_S", 51: "IQ1_S", 52: "IQ1_M", 53: "TQ1_0", 54: "TQ2_0", 55: "NVFP4",
	}
	if n, ok := names[dt]; ok {
		return n
	}
	return fmt.Sprintf("dtype(%d)", dt)
}

// formatSize is cmd/kuke/image/get.go's helper (1024-based, "-" for unknown); it would move to cmd/kuke/shared when a second user appears.
func formatSize(size int64) string {
	if size < 0 {
		return "-"
	}
	const unit = 1024
	if size < unit {
		return fmt.Sprintf("%d B", size)
	}
	div, exp := int64(unit), 0
	for n := size / unit; n >= unit; n /= unit {
		div *= unit
		exp++
	}
	return fmt.Sprintf("%.1f %ciB", float64(size)/float64(div), "KMGTPE"[exp])
}

var _ = gpupool.ModeSingle // keep the import when the table branch is compiled out in tests
